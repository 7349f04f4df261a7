"""DN-Splatter loss terms on fused CUDA kernels, behind the reference's class names.

`dn_regularizer_loss` evaluates the whole FusionSense depth / normal regulariser of dn_model.py:722-819
(/root/reference/dn_splatter/losses.py `EdgeAwareLogL1` :177-214, `TVLoss` :269-285, normal L1) in ONE fused
kernel pair (forward + analytic backward, csrc/losses.cu); `FusedSSIM` is the SSIM of the base splatfacto loss
(torchmetrics `StructuralSimilarityIndexMeasure(data_range=1.0, kernel_size=11)`, dn_model.py:244) in one launch each
way (csrc/ssim.cu).  `DepthLossType`, `DepthLoss`, `EdgeAwareLogL1`, `TVLoss` keep the reference's names and call
those kernels.  No torch arithmetic lives here: the plain-torch restatement that checks the kernels is
oracle/dn_losses_ref.py (test infrastructure).
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


class _FusedOnly(nn.Module):
    """Base of the class shells below: the reference's loss class names (dn_splatter/losses.py) on the fused CUDA
    kernel of csrc/losses.cu.  There is no torch arithmetic in this package; the plain-torch restatement that
    checks these kernels lives in oracle/dn_losses_ref.py (test infrastructure)."""

    @staticmethod
    def _refuse(what: str):
        raise NotImplementedError(
            f"{what} has no fused kernel in fusionsense_b200 (the FusionSense configuration does not use it, "
            "configs/config.py:3-39); use dn_splatter.losses from the reference for it")


class TVLoss(_FusedOnly):
    """losses.py:269-285: mean |d/dx| + mean |d/dy| of an [H,W,1] or [H,W,3] image, one fused launch."""

    def forward(self, pred: Tensor) -> Tensor:
        if pred.shape[-1] == 1:
            return dn_regularizer_loss(pred, None, None, None, None, sensor_depth_lambda=0.0, smooth_loss_lambda=1.0,
                                       normal_l1_lambda=0.0, normal_tv_lambda=0.0)
        if pred.shape[-1] == 3:
            return dn_regularizer_loss(None, None, None, pred, None, sensor_depth_lambda=0.0, smooth_loss_lambda=0.0,
                                       normal_l1_lambda=0.0, normal_tv_lambda=1.0)
        self._refuse(f"TVLoss on {pred.shape[-1]} channels")


class EdgeAwareLogL1(_FusedOnly):
    """losses.py:177-214 in its "scalar" form.  The valid mask is the one dn_model.py:724 builds,
    `gt > depth_tolerance`; it is evaluated inside the kernel, so `mask` may be omitted (a mask passed in must be
    that one)."""

    def __init__(self, implementation: Literal["scalar", "per-pixel"] = "scalar", depth_tolerance: float = 0.1,
                 rgb_clamp_min: float = 0.0, **kwargs):
        super().__init__()
        if implementation != "scalar":
            self._refuse("EdgeAwareLogL1(implementation='per-pixel')")
        self.depth_tolerance, self.rgb_clamp_min = float(depth_tolerance), float(rgb_clamp_min)

    def forward(self, pred: Tensor, gt: Tensor, rgb: Tensor, mask: Optional[Tensor] = None) -> Tensor:
        return dn_regularizer_loss(pred, gt, rgb, None, None, depth_tolerance=self.depth_tolerance,
                                   sensor_depth_lambda=1.0, smooth_loss_lambda=0.0, normal_l1_lambda=0.0,
                                   normal_tv_lambda=0.0, rgb_clamp_min=self.rgb_clamp_min)


class DepthLoss(_FusedOnly):
    """Factory with the reference's dispatch (losses.py:31-60) over the loss types that have a fused kernel."""

    def __init__(self, depth_loss_type: DepthLossType, **kwargs):
        super().__init__()
        self.depth_loss_type = depth_loss_type
        if depth_loss_type == DepthLossType.EdgeAwareLogL1:
            self.loss = EdgeAwareLogL1(**kwargs)
        elif depth_loss_type == DepthLossType.TV:
            self.loss = TVLoss()
        else:
            self._refuse(f"DepthLoss({depth_loss_type})")

    def forward(self, *args) -> Tensor:
        return self.loss(*args)


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
            # squeeze, not t[0]: select's backward materialises a zero [1,C,H,W] tensor and copies into it (two 25 MB
            # launches per 1080p step); squeeze's backward is a view.  Contiguous again when it came from an [H,W,C] image
            return t.squeeze(0).permute(1, 2, 0)
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
