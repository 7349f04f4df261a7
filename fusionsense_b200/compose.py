"""Fused image-space glue of `DNSplatterModel.get_outputs` and the flatness regulariser (csrc/compose.cu).

The reference issues these as ~20 torch launches forward and ~35 backward per iteration; at FusionSense's image
size every one of them is launch/latency-bound, and together they were a quarter of the captured step
(profiles/r01g_launches_graph.txt).  `dn_step.DNSplatterStep` uses them when `fused_outputs` is on; the
reference's own file keeps issuing the torch ops, which stay correct on the shim.

* `compose_rgbd`      — dn_model.py:602-604 + :609-613  (background blend, clamp, depth fill with the detached max)
* `normal_map`        — dn_model.py:655-656            ((n / |n| + 1) / 2)
* `flatness_loss`     — dn_model.py:817-819            (mean_i min_k exp(scales[i, k]))
* `combine_losses`    — dn_model.py:683-690, :925      (ssim_lambda (1 - ssim) + regulariser + normal_lambda flatness)
"""
from __future__ import annotations

from typing import Tuple

import torch
from torch import Tensor

from ._abi import check, lib, ptr
from .ops import _f32c, _req_cuda, _stream


class _ComposeRGBD(torch.autograd.Function):
    @staticmethod
    def forward(ctx, render, alpha, background):
        _req_cuda(render, alpha, background)
        render, alpha, background = _f32c(render), _f32c(alpha), _f32c(background.detach())
        assert render.shape[-1] == 4 and render.numel() == 4 * alpha.numel(), (render.shape, alpha.shape)
        assert background.numel() == 3, background.shape
        P = alpha.numel()
        H, W = render.shape[-3], render.shape[-2]
        assert P == H * W, "compose_rgbd handles one camera per call (dn_model.py:487)"
        dev = render.device
        rgb = torch.empty((H, W, 3), dtype=torch.float32, device=dev)
        depth = torch.empty((H, W, 1), dtype=torch.float32, device=dev)
        scratch = torch.empty((1,), dtype=torch.float32, device=dev)
        check(lib.fsb_compose_rgbd_fwd(P, ptr(render), ptr(alpha), ptr(background), ptr(rgb), ptr(depth), ptr(scratch),
                                       _stream()), "fsb_compose_rgbd_fwd")
        ctx.save_for_backward(render, alpha, background)
        ctx.set_materialize_grads(False)
        return rgb, depth

    @staticmethod
    def backward(ctx, v_rgb, v_depth):
        render, alpha, background = ctx.saved_tensors
        P = alpha.numel()
        v_render = torch.empty_like(render)
        v_alpha = torch.empty_like(alpha)
        check(lib.fsb_compose_rgbd_bwd(P, ptr(render), ptr(alpha), ptr(background), ptr(_f32c(v_rgb)),
                                       ptr(_f32c(v_depth)), ptr(v_render), ptr(v_alpha), _stream()),
              "fsb_compose_rgbd_bwd")
        return v_render, v_alpha, None


def compose_rgbd(render: Tensor, alpha: Tensor, background: Tensor) -> Tuple[Tensor, Tensor]:
    """render [1,H,W,4] (premultiplied RGB + expected depth), alpha [1,H,W,1], background [3] ->
    rgb [H,W,3] = clamp(render[..., :3] + (1 - alpha) * background, 0, 1),
    depth [H,W,1] = where(alpha > 0, render[..., 3:4], render[..., 3:4].detach().max())."""
    return _ComposeRGBD.apply(render, alpha, background)


class _NormalMap(torch.autograd.Function):
    @staticmethod
    def forward(ctx, normals_raw):
        _req_cuda(normals_raw)
        n = _f32c(normals_raw)
        assert n.shape[-1] == 3, n.shape
        out = torch.empty_like(n)
        check(lib.fsb_normal_map_fwd(n.numel() // 3, ptr(n), ptr(out), _stream()), "fsb_normal_map_fwd")
        ctx.save_for_backward(n)
        return out

    @staticmethod
    def backward(ctx, v_out):
        (n,) = ctx.saved_tensors
        v_raw = torch.empty_like(n)
        check(lib.fsb_normal_map_bwd(n.numel() // 3, ptr(n), ptr(_f32c(v_out)), ptr(v_raw), _stream()),
              "fsb_normal_map_bwd")
        return v_raw


def normal_map(normals_raw: Tensor) -> Tensor:
    """[..., 3] composited normals -> (n / |n| + 1) / 2, the [0, 1] map the losses and the viewer consume."""
    return _NormalMap.apply(normals_raw)


class _Flatness(torch.autograd.Function):
    @staticmethod
    def forward(ctx, log_scales):
        _req_cuda(log_scales)
        s = _f32c(log_scales)
        assert s.dim() == 2 and s.shape[1] == 3, s.shape
        ws = torch.empty((16,), dtype=torch.uint8, device=s.device)
        out = torch.empty((), dtype=torch.float32, device=s.device)
        check(lib.fsb_flatness_fwd(s.shape[0], ptr(s), ptr(ws), ptr(out), _stream()), "fsb_flatness_fwd")
        ctx.save_for_backward(s)
        return out

    @staticmethod
    def backward(ctx, v_out):
        (s,) = ctx.saved_tensors
        v = torch.empty_like(s)
        v_out = v_out.contiguous().float()
        check(lib.fsb_flatness_bwd(s.shape[0], ptr(s), ptr(v_out), ptr(v), _stream()), "fsb_flatness_bwd")
        return v


def flatness_loss(log_scales: Tensor) -> Tensor:
    """torch.min(torch.exp(log_scales), dim=1, keepdim=True)[0].mean() -> scalar tensor."""
    return _Flatness.apply(log_scales)


class _LossCombine(torch.autograd.Function):
    @staticmethod
    def forward(ctx, ssim, reg, flat, w_ssim, w_flat):
        ts = [None if t is None else _f32c(t) for t in (ssim, reg, flat)]
        ref = next(t for t in ts if t is not None)
        _req_cuda(ref)
        out = torch.empty((), dtype=torch.float32, device=ref.device)
        check(lib.fsb_loss_combine_fwd(ptr(ts[0]), ptr(ts[1]), ptr(ts[2]), float(w_ssim), float(w_flat), ptr(out),
                                       _stream()), "fsb_loss_combine_fwd")
        ctx.present = [t is not None for t in ts]
        ctx.w = (float(w_ssim), float(w_flat))
        return out

    @staticmethod
    def backward(ctx, v_out):
        v_out = v_out.contiguous().float()
        need = ctx.needs_input_grad
        vs = [torch.empty((), dtype=torch.float32, device=v_out.device) if (p and n) else None
              for p, n in zip(ctx.present, need[:3])]
        check(lib.fsb_loss_combine_bwd(ptr(v_out), ctx.w[0], ctx.w[1], ptr(vs[0]), ptr(vs[1]), ptr(vs[2]), _stream()),
              "fsb_loss_combine_bwd")
        return vs[0], vs[1], vs[2], None, None


def combine_losses(ssim, reg, flat, ssim_lambda: float, flat_lambda: float) -> Tensor:
    """ssim_lambda * (1 - ssim) + reg + flat_lambda * flat -> scalar tensor, in one launch each way (the reference
    assembles main_loss from ~7 scalar torch ops, dn_model.py:683-690 / :925).  Any term may be None."""
    return _LossCombine.apply(ssim, reg, flat, ssim_lambda, flat_lambda)


def u8_to_unit_float(src: Tensor, out: Tensor = None, recip: bool = True) -> Tensor:
    """uint8 CUDA tensor -> float32 in [0, 1], written into `out` (same element count), so the step's RGB / normal
    targets can cross PCIe as bytes.  `recip=True`: `src * (1.0f / 255.0f)`, the bits of `image.float() / 255.0` on a
    CUDA tensor (splatfacto's get_gt_img, SURVEY.md A.7: torch divides by a Python scalar through the fp32 reciprocal);
    `recip=False`: IEEE division, the bits of numpy's `normal_map.astype("float32") / 255.0` (dn_dataset.py:205)."""
    _req_cuda(src)
    assert src.dtype == torch.uint8 and src.is_contiguous(), (src.dtype, src.is_contiguous())
    if out is None:
        out = torch.empty(src.shape, dtype=torch.float32, device=src.device)
    assert out.dtype == torch.float32 and out.is_contiguous() and out.numel() == src.numel() and out.is_cuda
    check(lib.fsb_u8_to_unit_float(src.numel(), ptr(src), ptr(out), 1 if recip else 0, _stream()),
          "fsb_u8_to_unit_float")
    return out
