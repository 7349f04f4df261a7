"""Fused per-Gaussian helpers of the DN-Splatter step (csrc/gaussian_aux.cu).

* `gaussian_normals`  — dn_model.py:617-636 in one kernel (+ one for the backward).
* `densify_stats`     — splatfacto `after_train` accumulation in one kernel.
"""
from __future__ import annotations

from typing import Tuple

import torch
from torch import Tensor

from ._abi import check, lib, ptr
from .ops import _f32c, _req_cuda, _stream


class _GaussianNormals(torch.autograd.Function):
    @staticmethod
    def forward(ctx, quats, scales, means, c2w):
        _req_cuda(quats, scales, means, c2w)
        quats, scales, means = _f32c(quats.detach()), _f32c(scales.detach()), _f32c(means.detach())
        c2w = _f32c(c2w.detach().reshape(-1, 4)[:3])
        N = quats.shape[0]
        world = torch.empty((N, 3), dtype=torch.float32, device=quats.device)
        cam = torch.empty((N, 3), dtype=torch.float32, device=quats.device)
        check(lib.fsb_gaussian_normals_fwd(N, ptr(quats), ptr(scales), ptr(means), ptr(c2w), ptr(world), ptr(cam),
                                           _stream()), "fsb_gaussian_normals_fwd")
        ctx.save_for_backward(quats, scales, means, c2w)
        ctx.mark_non_differentiable(world)
        return cam, world

    @staticmethod
    def backward(ctx, v_cam, _v_world):
        quats, scales, means, c2w = ctx.saved_tensors
        N = quats.shape[0]
        v_quats = torch.empty_like(quats)
        check(lib.fsb_gaussian_normals_bwd(N, ptr(quats), ptr(scales), ptr(means), ptr(c2w), ptr(_f32c(v_cam)),
                                           ptr(v_quats), _stream()), "fsb_gaussian_normals_bwd")
        return v_quats, None, None, None


def gaussian_normals(quats: Tensor, scales: Tensor, means: Tensor, c2w: Tensor) -> Tuple[Tensor, Tensor]:
    """-> (normals in the camera frame [N,3] (differentiable w.r.t. quats), world-frame normals [N,3] (detached)).

    The Gaussian's normal is the rotation column of its smallest scale axis, flipped to face the camera
    (dn_model.py:617-633), then `normals @ c2w[:3,:3]` (:636).  c2w: [3,4] or [4,4] camera-to-world.
    """
    return _GaussianNormals.apply(quats, scales, means, c2w)


@torch.no_grad()
def densify_stats(radii: Tensor, grads2d: Tensor, max_dim: float, xys_grad_norm: Tensor, vis_counts: Tensor,
                  max_2Dsize: Tensor, skip_flag: Tensor = None) -> None:
    """In-place: for visible Gaussians (radii > 0) vis_counts += 1, xys_grad_norm += |grads2d|,
    max_2Dsize = max(max_2Dsize, radii / max_dim).  grads2d: `xys.absgrad[0]` (or `.grad[0]`), [N,2].
    `skip_flag` (device int32[1]): non-zero makes the call a no-op (overflowed step of a captured graph)."""
    _req_cuda(radii, grads2d, xys_grad_norm, vis_counts, max_2Dsize)
    N = radii.numel()
    assert radii.dtype == torch.int32 and radii.is_contiguous()
    g = _f32c(grads2d)
    for t in (xys_grad_norm, vis_counts, max_2Dsize):
        assert t.dtype == torch.float32 and t.is_contiguous() and t.numel() == N
    check(lib.fsb_densify_stats(N, ptr(radii), ptr(g), float(max_dim), ptr(xys_grad_norm), ptr(vis_counts),
                                ptr(max_2Dsize), ptr(skip_flag), _stream()), "fsb_densify_stats")
