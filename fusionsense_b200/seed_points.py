"""Seed point cloud of Module 1 on the GPU (csrc/seed_points.cu) — SURVEY.md §8 f3.

Host mirror of `utils/generate_pcd.py`:

* `get_pointcloud(color, depth, w2c, FX, FY, CX, CY, transform_pts=True, mask=None)` — generate_pcd.py:15-48, same
  arguments and return value: `(fore_pcd, back_pcd)`, float32 `[n, 6]` rows (world xyz, rgb) of the pixels with
  0 < depth < 0.5 and 0.5 < depth < 5, in pixel order.  (`mask` is accepted and ignored, as in the reference.)
* `voxel_down_sample(rows, voxel_size)` — open3d's `PointCloud.voxel_down_sample` as generate_pcd.py:98-101 uses it on
  the per-view background cloud: one row per occupied voxel, the fp64 mean of the voxel's points and colours.
  open3d returns the voxels in its hash map's iteration order; here they come in ascending (x, y, z) voxel order.
* `merged_background_cloud(views, ...)` — the loop of `init_pcd_generate` (generate_pcd.py:86-101) without the file
  I/O: per view back-project, down-sample, concatenate.

No CPU path: CPU tensors raise.
"""
from __future__ import annotations

import ctypes
from typing import Iterable, Optional, Sequence, Tuple

import torch
from torch import Tensor

from ._abi import FsbError, check, lib, ptr
from .ops import _f32c, _req_cuda, _stream, isect_scan, radix_sort_pairs


def _scalar(v) -> float:
    return float(v.item()) if isinstance(v, Tensor) else float(v)


def _backproject(color: Tensor, depth: Tensor, rot, trans, fx, fy, cx, cy, lo: float, hi: float) -> Tensor:
    H, W = depth.shape[-2], depth.shape[-1]
    P = H * W
    dev = depth.device
    flags = torch.empty((P,), dtype=torch.int32, device=dev)
    check(lib.fsb_backproject_flags(P, ptr(depth), lo, hi, ptr(flags), _stream()), "fsb_backproject_flags")
    offsets, n = isect_scan(flags)
    out = torch.empty((int(n), 6), dtype=torch.float32, device=dev)
    if int(n) == 0:
        return out
    r = (ctypes.c_float * 9)(*rot)
    t = (ctypes.c_float * 3)(*trans)
    check(lib.fsb_backproject_emit(H, W, ptr(depth), ptr(color), ctypes.addressof(r), ctypes.addressof(t), fx, fy, cx, cy,
                                   lo, hi, ptr(offsets), ptr(out), _stream()), "fsb_backproject_emit")
    return out


def get_pointcloud(color: Tensor, depth: Tensor, w2c: Tensor, FX, FY, CX, CY, transform_pts: bool = True,
                   mask: Optional[Tensor] = None) -> Tuple[Tensor, Tensor]:
    """utils/generate_pcd.py:15-48.  color [3,H,W] (ToTensor()), depth [H,W] in metres, w2c [4,4]."""
    _req_cuda(color, depth)
    color, depth = _f32c(color), _f32c(depth)
    assert color.dim() == 3 and color.shape[0] == 3 and depth.shape == color.shape[1:], (color.shape, depth.shape)
    if transform_pts:
        c2w = torch.inverse(w2c.float())  # generate_pcd.py:30, on the device the caller keeps w2c on
        c2w = c2w.cpu()
        rot = [float(v) for v in c2w[:3, :3].reshape(-1)]
        trans = [float(v) for v in c2w[:3, 3]]
    else:
        rot, trans = [1.0, 0, 0, 0, 1.0, 0, 0, 0, 1.0], [0.0, 0.0, 0.0]
    fx, fy, cx, cy = _scalar(FX), _scalar(FY), _scalar(CX), _scalar(CY)
    fore = _backproject(color, depth, rot, trans, fx, fy, cx, cy, 0.0, 0.5)   # generate_pcd.py:42
    back = _backproject(color, depth, rot, trans, fx, fy, cx, cy, 0.5, 5.0)   # generate_pcd.py:43
    return fore, back


def voxel_down_sample(rows: Tensor, voxel_size: float) -> Tensor:
    """rows [N, >=3] float32 (xyz first; every column is averaged, at most 9) -> float64 [n_voxels, width]."""
    _req_cuda(rows)
    rows = _f32c(rows)
    assert rows.dim() == 2 and 3 <= rows.shape[1] <= 9, rows.shape
    if voxel_size <= 0.0:
        raise ValueError("voxel_size <= 0.")  # open3d's message
    n, width = rows.shape
    dev = rows.device
    if n == 0:
        return torch.empty((0, width), dtype=torch.float64, device=dev)
    keys = torch.empty((n,), dtype=torch.int64, device=dev)
    vals = torch.empty((n,), dtype=torch.int32, device=dev)
    min_bound = torch.empty((3,), dtype=torch.float64, device=dev)
    overflow = torch.zeros((1,), dtype=torch.int32, device=dev)
    ws_bytes = lib.fsb_voxel_workspace()
    ws = torch.empty((ws_bytes,), dtype=torch.uint8, device=dev)
    check(lib.fsb_voxel_keys(n, ptr(rows), width, float(voxel_size), ptr(min_bound), ptr(keys), ptr(vals), ptr(overflow),
                             ptr(ws), ws_bytes, _stream()), "fsb_voxel_keys")
    keys, vals = radix_sort_pairs(keys, vals, 63)
    heads = torch.empty((n,), dtype=torch.int32, device=dev)
    check(lib.fsb_voxel_heads(n, ptr(keys), ptr(heads), _stream()), "fsb_voxel_heads")
    offsets, m = isect_scan(heads)
    if int(overflow) != 0:
        raise FsbError("voxel_size is too small.")  # open3d's refusal (index range), here at 2^21 voxels per axis
    out = torch.empty((int(m), width), dtype=torch.float64, device=dev)
    check(lib.fsb_voxel_mean(n, ptr(keys), ptr(vals), ptr(heads), ptr(offsets), ptr(rows), width, width, ptr(out),
                             _stream()), "fsb_voxel_mean")
    return out


def merged_background_cloud(views: Iterable[Tuple[Tensor, Tensor, Tensor]], FX, FY, CX, CY,
                            voxel_size: float = 0.02) -> Tensor:
    """generate_pcd.py:86-101: for every (color [3,H,W], depth [H,W], w2c [4,4]) view the background points
    (0.5 < depth < 5) are voxel-down-sampled and appended.  Returns float64 [M, 6] (xyz, rgb)."""
    parts = []
    for color, depth, w2c in views:
        _, back = get_pointcloud(color, depth, w2c, FX, FY, CX, CY, transform_pts=True)
        parts.append(voxel_down_sample(back, voxel_size))
    if not parts:
        raise ValueError("no views")
    return torch.cat(parts, dim=0)
