"""DN-Splatter loss terms behind the reference's class names.

Mirrors /root/reference/dn_splatter/losses.py (`DepthLossType`, `DepthLoss`, `LogL1` :161-174,
`EdgeAwareLogL1` :177-214, `EdgeAwareTV` :241-266, `TVLoss` :269-285) and the SSIM the base splatfacto loss
uses (torchmetrics `StructuralSimilarityIndexMeasure(data_range=1.0, kernel_size=11)`, dn_model.py:244).
`dn_regularizer_loss` evaluates the whole FusionSense depth/normal regulariser of
dn_model.py:722-819 in ONE fused CUDA kernel pair (forward + analytic backward), see csrc/losses.cu.
"""
from __future__ import annotations

from enum import Enum
from typing import Literal, Optional

import torch
import torch.nn.functional as F
from torch import Tensor, nn


class DepthLossType(Enum):
    MSE = "mse"
    L1 = "L1"
    LogL1 = "LogL1"
    HuberL1 = "HuberL1"
    TV = "TV"
    EdgeAwareLogL1 = "EdgeAwareLogL1"
    EdgeAwareTV = "EdgeAwareTV"


class LogL1(nn.Module):
    def __init__(self, implementation: Literal["scalar", "per-pixel"] = "scalar", **kwargs):
        super().__init__()
        self.implementation = implementation

    def forward(self, pred, gt):
        v = torch.log(1 + torch.abs(pred - gt))
        return v.mean() if self.implementation == "scalar" else v


class L1(nn.Module):
    def __init__(self, implementation: Literal["scalar", "per-pixel"] = "scalar", **kwargs):
        super().__init__()
        self.implementation = implementation

    def forward(self, pred, gt):
        v = torch.abs(pred - gt)
        return v.mean() if self.implementation == "scalar" else v


class EdgeAwareLogL1(nn.Module):
    """log(1+|d - d_gt|) weighted by exp(-mean_c |grad rgb|) in x and y, masked means (losses.py:177-214)."""

    def __init__(self, implementation: Literal["scalar", "per-pixel"] = "scalar", **kwargs):
        super().__init__()
        self.implementation = implementation
        self.logl1 = LogL1(implementation="per-pixel")

    def forward(self, pred: Tensor, gt: Tensor, rgb: Tensor, mask: Optional[Tensor]):
        logl1 = self.logl1(pred, gt)
        grad_img_x = torch.mean(torch.abs(rgb[..., :, :-1, :] - rgb[..., :, 1:, :]), -1, keepdim=True)
        grad_img_y = torch.mean(torch.abs(rgb[..., :-1, :, :] - rgb[..., 1:, :, :]), -1, keepdim=True)
        lambda_x = torch.exp(-grad_img_x)
        lambda_y = torch.exp(-grad_img_y)
        loss_x = lambda_x * logl1[..., :, :-1, :]
        loss_y = lambda_y * logl1[..., :-1, :, :]
        if self.implementation == "per-pixel":
            if mask is not None:
                loss_x[~mask[..., :, :-1, :]] = 0
                loss_y[~mask[..., :-1, :, :]] = 0
            return loss_x[..., :-1, :, :] + loss_y[..., :, :-1, :]
        if mask is not None:
            assert mask.shape[:2] == pred.shape[:2]
            loss_x = loss_x[mask[..., :, :-1, :]]
            loss_y = loss_y[mask[..., :-1, :, :]]
        return loss_x.mean() + loss_y.mean()


class EdgeAwareTV(nn.Module):
    def forward(self, depth: Tensor, rgb: Tensor):
        grad_depth_x = torch.abs(depth[..., :, :-1, :] - depth[..., :, 1:, :])
        grad_depth_y = torch.abs(depth[..., :-1, :, :] - depth[..., 1:, :, :])
        grad_img_x = torch.mean(torch.abs(rgb[..., :, :-1, :] - rgb[..., :, 1:, :]), -1, keepdim=True)
        grad_img_y = torch.mean(torch.abs(rgb[..., :-1, :, :] - rgb[..., 1:, :, :]), -1, keepdim=True)
        grad_depth_x = grad_depth_x * torch.exp(-grad_img_x)
        grad_depth_y = grad_depth_y * torch.exp(-grad_img_y)
        return grad_depth_x.mean() + grad_depth_y.mean()


class TVLoss(nn.Module):
    def forward(self, pred):
        h_diff = pred[..., :, :-1, :] - pred[..., :, 1:, :]
        w_diff = pred[..., :-1, :, :] - pred[..., 1:, :, :]
        return torch.mean(torch.abs(h_diff)) + torch.mean(torch.abs(w_diff))


class DepthLoss(nn.Module):
    """Factory with the reference's dispatch (losses.py:31-60)."""

    def __init__(self, depth_loss_type: DepthLossType, **kwargs):
        super().__init__()
        self.depth_loss_type = depth_loss_type
        self.kwargs = kwargs
        t = depth_loss_type
        if t == DepthLossType.MSE:
            self.loss = torch.nn.MSELoss()
        elif t == DepthLossType.L1:
            self.loss = L1(**kwargs)
        elif t == DepthLossType.LogL1:
            self.loss = LogL1(**kwargs)
        elif t == DepthLossType.EdgeAwareLogL1:
            self.loss = EdgeAwareLogL1(**kwargs)
        elif t == DepthLossType.EdgeAwareTV:
            self.loss = EdgeAwareTV()
        elif t == DepthLossType.TV:
            self.loss = TVLoss()
        else:
            raise ValueError(f"Unsupported loss type: {depth_loss_type}")

    def forward(self, *args) -> Tensor:
        return self.loss(*args)


# ---------------------------------------------------------------------------------------------
# SSIM as torchmetrics' StructuralSimilarityIndexMeasure computes it (gaussian 11x11, sigma 1.5,
# reflect padding, border crop, mean).  Plain-torch restatement: the checker of FusedSSIM below and what the CPU
# reference arm runs; the CUDA product path uses FusedSSIM.
# ---------------------------------------------------------------------------------------------
class SSIM(nn.Module):
    def __init__(self, data_range: float = 1.0, kernel_size: int = 11, sigma: float = 1.5, k1=0.01, k2=0.03):
        super().__init__()
        self.data_range, self.kernel_size, self.k1, self.k2 = data_range, kernel_size, k1, k2
        dist = torch.arange((1 - kernel_size) / 2, (1 + kernel_size) / 2, 1.0)
        g = torch.exp(-((dist / sigma) ** 2) / 2)
        g = (g / g.sum())[None]
        self.register_buffer("kernel2d", (g.t() @ g)[None, None], persistent=False)

    def forward(self, preds: Tensor, target: Tensor) -> Tensor:  # [B,C,H,W]
        c1, c2 = (self.k1 * self.data_range) ** 2, (self.k2 * self.data_range) ** 2
        ch = preds.shape[1]
        pad = (self.kernel_size - 1) // 2
        kernel = self.kernel2d.to(preds.dtype).expand(ch, 1, -1, -1)
        p = F.pad(preds, (pad, pad, pad, pad), mode="reflect")
        t = F.pad(target, (pad, pad, pad, pad), mode="reflect")
        x = torch.cat((p, t, p * p, t * t, p * t))
        out = F.conv2d(x, kernel, groups=ch)
        mu_p, mu_t, e_pp, e_tt, e_pt = out.split(preds.shape[0])
        mu_pp, mu_tt, mu_pt = mu_p * mu_p, mu_t * mu_t, mu_p * mu_t
        s_p, s_t, s_pt = e_pp - mu_pp, e_tt - mu_tt, e_pt - mu_pt
        ssim = ((2 * mu_pt + c1) * (2 * s_pt + c2)) / ((mu_pp + mu_tt + c1) * (s_p + s_t + c2))
        return ssim[..., pad:-pad, pad:-pad].reshape(ssim.shape[0], -1).mean(-1).mean()


class _FusedSSIMFn(torch.autograd.Function):
    """ssim(x, y) with the gradient taken w.r.t. x (SSIM is symmetric, the caller orders the arguments)."""

    @staticmethod
    def forward(ctx, x, y, taps, data_range, k1, k2):
        from ._abi import check, lib, ptr
        from .ops import _f32c, _stream

        if not (x.is_cuda and y.is_cuda):
            raise RuntimeError("FusedSSIM needs CUDA tensors (no CPU fallback)")
        x, y = _f32c(x), _f32c(y)
        H, W, C = x.shape
        need = ctx.needs_input_grad[0]
        dev = x.device
        ws = torch.empty((lib.fsb_ssim_workspace(),), dtype=torch.uint8, device=dev)
        out = torch.empty((), dtype=torch.float32, device=dev)
        maps = torch.empty((3, H - 10, W - 10, C), dtype=torch.float32, device=dev) if need else None
        d = [maps[i] for i in range(3)] if need else [None] * 3
        from .ops import kernel_timer
        ev = kernel_timer.start("ssim_fwd")
        check(lib.fsb_ssim_fwd(H, W, C, ptr(x), ptr(y), data_range, k1, k2, taps, ptr(ws), ptr(out), ptr(d[0]),
                               ptr(d[1]), ptr(d[2]), _stream()), "fsb_ssim_fwd")
        kernel_timer.stop(ev)
        if need:
            ctx.save_for_backward(x, y, maps)
        ctx.cfg = (taps, data_range, k1, k2)
        return out

    @staticmethod
    def backward(ctx, v_out):
        from ._abi import check, lib, ptr
        from .ops import _stream

        x, y, maps = ctx.saved_tensors
        taps, data_range, k1, k2 = ctx.cfg
        H, W, C = x.shape
        v_x = torch.empty_like(x)
        v_out = v_out.contiguous().float()
        from .ops import kernel_timer
        ev = kernel_timer.start("ssim_bwd")
        check(lib.fsb_ssim_bwd(H, W, C, ptr(x), ptr(y), data_range, k1, k2, taps, ptr(maps[0]), ptr(maps[1]),
                               ptr(maps[2]), ptr(v_out), ptr(v_x), _stream()), "fsb_ssim_bwd")
        kernel_timer.stop(ev)
        return v_x, None, None, None, None, None


class FusedSSIM(nn.Module):
    """Drop-in for the `self.ssim` object of dn_model.py:244 on libfsb200's kernels (csrc/ssim.cu).

    Called like torchmetrics: `ssim(preds[B,C,H,W], target[B,C,H,W]) -> scalar` with B == 1 (one camera per step,
    dn_model.py:487); also accepts [H,W,C] images directly, which skips the permute copies.  One launch forward,
    one backward, against ~45 torch launches each way.
    """

    def __init__(self, data_range: float = 1.0, kernel_size: int = 11, sigma: float = 1.5, k1=0.01, k2=0.03):
        super().__init__()
        import ctypes

        if kernel_size != 11:
            raise ValueError("FusedSSIM implements the reference's setting only: kernel_size=11")
        self.data_range, self.k1, self.k2 = float(data_range), float(k1), float(k2)
        dist = torch.arange((1 - kernel_size) / 2, (1 + kernel_size) / 2, 1.0)
        g = torch.exp(-((dist / sigma) ** 2) / 2)
        g = g / g.sum()
        self._taps_arr = (ctypes.c_float * kernel_size)(*[float(v) for v in g])
        self._taps = ctypes.addressof(self._taps_arr)

    @staticmethod
    def _hwc(t: Tensor) -> Tensor:
        if t.dim() == 4:
            if t.shape[0] != 1:
                raise ValueError("FusedSSIM handles one image per call (the reference trains one camera per step)")
            return t[0].permute(1, 2, 0)  # a view; contiguous again when it came from an [H,W,C] image
        return t

    def forward(self, preds: Tensor, target: Tensor) -> Tensor:
        p, t = self._hwc(preds), self._hwc(target)
        if t.requires_grad and not p.requires_grad:
            return _FusedSSIMFn.apply(t, p, self._taps, self.data_range, self.k1, self.k2)
        if p.requires_grad and t.requires_grad:
            # both sides differentiable: two symmetric evaluations, each carrying one gradient
            a = _FusedSSIMFn.apply(p, t.detach(), self._taps, self.data_range, self.k1, self.k2)
            b = _FusedSSIMFn.apply(t, p.detach(), self._taps, self.data_range, self.k1, self.k2)
            return a + (b - b.detach())
        return _FusedSSIMFn.apply(p, t, self._taps, self.data_range, self.k1, self.k2)


# ---------------------------------------------------------------------------------------------
# fused CUDA path (csrc/losses.cu): the FusionSense regulariser in one forward + one backward launch
# ---------------------------------------------------------------------------------------------
class _DNRegulariser(torch.autograd.Function):
    @staticmethod
    def forward(ctx, depth, pred_normal, pred_rgb, sensor, edge_rgb, gt_normal, gt_rgb, depth_tol, rgb_clamp_min,
                l_sensor, l_smooth, l_nl1, l_ntv, l_rgb):
        from ._abi import check, lib, ptr
        from .ops import _f32c, _stream

        ref = depth if depth is not None else (pred_normal if pred_normal is not None else pred_rgb)
        if not ref.is_cuda:
            raise RuntimeError("dn_regularizer_loss needs CUDA tensors (no CPU fallback)")
        H, W = ref.shape[0], ref.shape[1]
        ts = [_f32c(t) for t in (depth, sensor, edge_rgb, pred_normal, gt_normal, pred_rgb, gt_rgb)]
        ws = torch.empty((lib.fsb_dn_loss_workspace(),), dtype=torch.uint8, device=ref.device)
        loss = torch.empty((), dtype=torch.float32, device=ref.device)
        weights = (float(depth_tol), float(rgb_clamp_min), float(l_sensor), float(l_smooth), float(l_nl1),
                   float(l_ntv), float(l_rgb))
        check(lib.fsb_dn_loss_fwd(H, W, *[ptr(t) for t in ts], *weights, ptr(ws), ptr(loss), _stream()),
              "fsb_dn_loss_fwd")
        ctx.save_for_backward(*[t for t in ts if t is not None], ws)
        ctx.present = [t is not None for t in ts]
        ctx.cfg = (H, W, weights)
        return loss

    @staticmethod
    def backward(ctx, v_loss):
        from ._abi import check, lib, ptr
        from .ops import _stream

        saved = list(ctx.saved_tensors)
        ws = saved.pop()
        it = iter(saved)
        ts = [next(it) if p else None for p in ctx.present]
        depth, sensor, edge_rgb, pred_normal, gt_normal, pred_rgb, gt_rgb = ts
        H, W, weights = ctx.cfg
        need = ctx.needs_input_grad
        v_depth = torch.empty_like(depth) if (depth is not None and need[0]) else None
        v_normal = torch.empty_like(pred_normal) if (pred_normal is not None and need[1]) else None
        v_rgb = torch.empty_like(pred_rgb) if (pred_rgb is not None and need[2]) else None
        v_loss = v_loss.contiguous().float()
        check(lib.fsb_dn_loss_bwd(H, W, *[ptr(t) for t in ts], *weights, ptr(ws), ptr(v_loss), ptr(v_depth),
                                  ptr(v_normal), ptr(v_rgb), _stream()), "fsb_dn_loss_bwd")
        return (v_depth, v_normal, v_rgb) + (None,) * 11


def dn_regularizer_loss(depth_out: Optional[Tensor], sensor_depth_gt: Optional[Tensor], gt_img: Optional[Tensor],
                        pred_normal: Optional[Tensor], gt_normal: Optional[Tensor], pred_rgb: Optional[Tensor] = None,
                        gt_rgb: Optional[Tensor] = None, depth_tolerance: float = 0.1, sensor_depth_lambda: float = 0.2,
                        smooth_loss_lambda: float = 0.1, normal_l1_lambda: float = 0.4, normal_tv_lambda: float = 0.4,
                        rgb_l1_lambda: float = 0.0, rgb_clamp_min: float = 10 / 255.0) -> Tensor:
    """sensor_depth_lambda * EdgeAwareLogL1(depth_out, sensor_depth_gt, gt_img.clamp(min), sensor > tol)
    + smooth_loss_lambda * TV(depth_out) + normal_l1_lambda * |gt_normal - pred_normal|.mean()
    + normal_tv_lambda * TV(pred_normal) + rgb_l1_lambda * |gt_rgb - pred_rgb|.mean()   -> scalar tensor.

    The terms of dn_model.py:722-736, :753-756, :806, :814-815 (each already multiplied by its lambda, the normal
    terms by normal_lambda) and the L1 half of the base splatfacto loss, evaluated by one fused kernel.
    Image shapes [H,W,1] / [H,W,3]; differentiable w.r.t. depth_out, pred_normal, pred_rgb.
    """
    return _DNRegulariser.apply(depth_out, pred_normal, pred_rgb, sensor_depth_gt, gt_img, gt_normal, gt_rgb,
                                depth_tolerance, rgb_clamp_min, sensor_depth_lambda, smooth_loss_lambda,
                                normal_l1_lambda, normal_tv_lambda, rgb_l1_lambda)
