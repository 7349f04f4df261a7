"""`dn_splatter.utils.normal_utils` on the B200 kernels (csrc/pseudo_normals.cu).

Mirrors /root/reference/dn_splatter/utils/normal_utils.py: same function names, argument order and meaning.
`normal_from_depth_image` is what dn_model.py:779-789 calls when `normal_supervision == "depth"`.
There is no CPU path: CPU tensors raise.
"""
from __future__ import annotations

import ctypes

import torch
from torch import Tensor

from .._abi import check, lib, ptr
from ..ops import _f32c, _req_cuda, _stream


def _host3(vals, n):
    arr = (ctypes.c_float * n)(*[float(v) for v in vals])
    return arr


def pcd_to_normal(xyz: Tensor) -> Tensor:
    """normal_utils.py:7-20 — xyz [H,W,3] -> normals [H,W,3] (cross of central differences, zero border)."""
    _req_cuda(xyz)
    if xyz.dim() != 3 or xyz.shape[-1] != 3:
        raise ValueError(f"pcd_to_normal expects [H,W,3], got {tuple(xyz.shape)}")
    xyz = _f32c(xyz.detach())
    H, W, _ = xyz.shape
    out = torch.empty((H, W, 3), dtype=torch.float32, device=xyz.device)
    check(lib.fsb_normal_from_depth(H, W, None, ptr(xyz), 1.0, 1.0, 0.0, 0.0, None, None, ptr(out), _stream()),
          "fsb_normal_from_depth")
    return out


def normal_from_depth_image(depths: Tensor, fx: float, fy: float, cx: float, cy: float, img_size: tuple, c2w: Tensor,
                            device: torch.device, smooth: bool = False) -> Tensor:
    """estimate normals from depth map (normal_utils.py:23-46).  img_size is (W, H) as in the reference.

    `smooth=True`: the reference prints a notice and skips the filter when the depth map has any non-zero
    element (normal_utils.py:36-37); its other branch references an un-imported `cv2` and cannot run, so it is
    refused here as well.
    """
    _req_cuda(depths)
    if smooth:
        if torch.count_nonzero(depths) > 0:
            print("Input depth map contains 0 elements, skipping smoothing filter")
        else:
            raise NameError("name 'cv2' is not defined")  # what the reference raises on this branch
    W, H = int(img_size[0]), int(img_size[1])
    d = _f32c(depths.detach()).reshape(-1)
    if d.numel() != H * W:
        raise RuntimeError(f"depths has {d.numel()} elements, img_size {img_size} needs {H * W}")
    rot_inv = trans = None
    if c2w is not None:
        c = c2w.detach().float().cpu()
        rot_inv = _host3(torch.linalg.inv(c[..., :3, :3]).reshape(-1).tolist(), 9)
        trans = _host3(c[..., :3, 3].reshape(-1).tolist(), 3)
    out = torch.empty((H, W, 3), dtype=torch.float32, device=d.device)
    check(lib.fsb_normal_from_depth(H, W, ptr(d), None, float(fx), float(fy), float(cx), float(cy), rot_inv, trans,
                                    ptr(out), _stream()), "fsb_normal_from_depth")
    return out
