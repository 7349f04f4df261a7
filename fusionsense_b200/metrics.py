"""Per-step image metrics on fused kernels, behind the reference's class names.

`DNSplatterModel.get_metrics_dict` (/root/reference/dn_splatter/dn_model.py:927-1000) runs on EVERY training iteration:
`RGBMetrics` (dn_splatter/metrics.py:77-108: torchmetrics PSNR, SSIM and an LPIPS CNN), `MSELoss`, `DepthMetrics`
(:111-150) and eleven `float()` conversions, each a host synchronisation.  Here:

  * `image_metrics()`   PSNR / MSE / the seven depth metrics in ONE launch (csrc/metrics.cu), left on the device;
  * `RGBMetrics`        same call signature and return tuple as the reference's class; SSIM through `FusedSSIM`
                        (csrc/ssim.cu); LPIPS is a user-supplied callable (no backbone weights ship with this package)
                        evaluated every `lpips_every` calls — in between the last value is returned;
  * `DepthMetrics`      same signature, fused.

Nothing here synchronises with the host; a caller that wants Python floats converts when it logs."""
from __future__ import annotations

from typing import Callable, Dict, Optional

import torch
from torch import Tensor, nn

from ._abi import check, lib, ptr
from .losses import FusedSSIM
from .ops import _f32c, _stream

NAMES = ("rgb_mse", "rgb_psnr", "depth_abs_rel", "depth_sq_rel", "depth_rmse", "depth_rmse_log", "depth_a1", "depth_a2",
         "depth_a3", "depth_n_valid")
_WS: Dict[torch.device, Tensor] = {}


def _workspace(device) -> Tensor:
    ws = _WS.get(device)
    if ws is None:
        ws = torch.zeros((lib.fsb_image_metrics_workspace(),), dtype=torch.uint8, device=device)
        _WS[device] = ws
    return ws


@torch.no_grad()
def image_metrics(pred_rgb: Optional[Tensor], gt_rgb: Optional[Tensor], pred_depth: Optional[Tensor] = None,
                  gt_depth: Optional[Tensor] = None, depth_tolerance: float = 0.1, out: Optional[Tensor] = None) -> Tensor:
    """[H,W,3] images and / or [H,W(,1)] depths -> float32 [10] on the device, in the order of `NAMES`."""
    ref = pred_rgb if pred_rgb is not None else pred_depth
    if ref is None or not ref.is_cuda:
        raise RuntimeError("fusionsense_b200.metrics needs CUDA tensors (no CPU fallback)")
    H, W = ref.shape[0], ref.shape[1]
    pr, gr, pd, gd = (_f32c(t) for t in (pred_rgb, gt_rgb, pred_depth, gt_depth))
    if out is None:
        out = torch.empty((10,), dtype=torch.float32, device=ref.device)
    ws = _workspace(ref.device)
    check(lib.fsb_image_metrics(H, W, ptr(pr), ptr(gr), ptr(pd), ptr(gd), float(depth_tolerance), ptr(ws), ptr(out),
                                _stream()), "fsb_image_metrics")
    return out


class DepthMetrics(nn.Module):
    """dn_splatter/metrics.py:111-150: forward(pred, gt) -> (abs_rel, sq_rel, rmse, rmse_log, a1, a2, a3), device scalars."""

    def __init__(self, tolerance: float = 0.1, **kwargs):
        super().__init__()
        self.tolerance = tolerance

    @torch.no_grad()
    def forward(self, pred: Tensor, gt: Tensor):
        hw = lambda t: t.reshape(t.shape[-2], t.shape[-1]) if t.dim() == 3 and t.shape[0] == 1 else t.squeeze(-1)  # noqa: E731
        m = image_metrics(None, None, hw(pred), hw(gt), self.tolerance)
        return tuple(m[i] for i in range(2, 9))


class RGBMetrics(nn.Module):
    """dn_splatter/metrics.py:77-108: forward(pred[B,C,H,W], gt[B,C,H,W]) -> (psnr, ssim, lpips), device scalars.

    `lpips`: callable (pred, gt) -> scalar tensor, e.g. torchmetrics' LearnedPerceptualImagePatchSimilarity where it is
    installed; evaluated every `lpips_every` calls (0 / None callable: never, NaN is returned)."""

    def __init__(self, lpips: Optional[Callable] = None, lpips_every: int = 100, **kwargs):
        super().__init__()
        self.ssim = FusedSSIM(data_range=1.0, kernel_size=11)
        self.lpips, self.lpips_every = lpips, int(lpips_every)
        self._calls = 0
        self._last_lpips: Optional[Tensor] = None

    @torch.no_grad()
    def forward(self, pred: Tensor, gt: Tensor):
        if pred.dim() != 4 or pred.shape[0] != 1:
            raise ValueError("RGBMetrics handles one image per call ([1,C,H,W], as dn_model.py:962-966 passes it)")
        p, g = pred[0].permute(1, 2, 0), gt[0].permute(1, 2, 0)
        m = image_metrics(p, g)
        ssim = self.ssim(p.detach(), g.detach())
        if self._last_lpips is None:
            self._last_lpips = torch.full((), float("nan"), device=pred.device)
        if self.lpips is not None and self.lpips_every > 0 and self._calls % self.lpips_every == 0:
            self._last_lpips = self.lpips(pred, gt).detach()
        self._calls += 1
        return m[1], ssim, self._last_lpips


@torch.no_grad()
def step_metrics(outputs: Dict[str, Tensor], batch: Dict[str, Tensor], depth_tolerance: float = 0.1,
                 ssim: Optional[FusedSSIM] = None, out: Optional[Tensor] = None) -> Dict[str, Tensor]:
    """get_metrics_dict of dn_model.py:927-1000 on device tensors: one metrics launch (+ one SSIM launch)."""
    depth = outputs.get("depth")
    m = image_metrics(outputs["rgb"], batch["image"], depth.squeeze(-1) if depth is not None else None,
                      batch["sensor_depth"].squeeze(-1) if "sensor_depth" in batch else None, depth_tolerance, out=out)
    d = {n: m[i] for i, n in enumerate(NAMES)}
    if ssim is not None:
        d["rgb_ssim"] = ssim(outputs["rgb"].detach(), batch["image"])
    return d
