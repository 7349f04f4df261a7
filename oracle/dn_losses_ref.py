"""Plain-torch restatement of the reference's loss classes (and of the torchmetrics SSIM its base loss uses).

TEST INFRASTRUCTURE ONLY: nothing under fusionsense_b200/ imports this module; tests/, __graft_entry__.smoke() and
bench.py's CPU legs hand it to `DNSplatterStep(torch_losses=...)` as the checker / CPU arm of the fused kernels.

Follows /root/reference/dn_splatter/losses.py (`DepthLossType` :18-28, `DepthLoss` :31-60, `L1` / `LogL1` :145-174,
`EdgeAwareLogL1` :177-214, `EdgeAwareTV` :241-266, `TVLoss` :269-285) and torchmetrics'
`StructuralSimilarityIndexMeasure(data_range=1.0, kernel_size=11)` as dn_model.py:244 constructs it.
PINNED: tests/golden/dn_losses.npz holds outputs and gradients of the unmodified reference classes
(oracle/make_golden_losses.py); tests/test_losses_vs_reference_golden.py checks this file against them.
"""
from __future__ import annotations

from typing import Literal, Optional

import torch
import torch.nn.functional as F
from torch import Tensor, nn

from fusionsense_b200.losses import DepthLossType  # the enum only (shared names, no arithmetic)


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
