"""Level-set surface points for the mesh export — host mirror of
`DNSplatterModel.compute_level_surface_points` (dn_splatter/dn_model.py:1705-1946, called per training camera by
`gs-mesh sugar-coarse`, dn_splatter/export_mesh.py:543).

The reference renders the view, back-projects the depth map, looks up the 16 nearest Gaussians of every point with
sklearn on the CPU, samples 21 points along every camera ray, evaluates the density of the 16 neighbours at each sample
through `[P*21, 16, 3, 3]` temporaries in passes of 2M samples, and searches every ray for the first crossing of each
surface level.  Here: `fusionsense_b200.knn.knn_sk` on the device and ONE kernel (`fsb_level_crossings`) for
std + sampling + densities + crossing search; the glue around them (render, back-projection, gathers, the random
subsample) stays torch, statement for statement.

`level_crossings(...)` is the kernel's binding; `compute_level_surface_points(model, camera, ...)` takes the model as
its first argument, so `DNSplatterModel.compute_level_surface_points = fusionsense_b200.level_set.compute_level_surface_points`
is the whole patch.
"""
from __future__ import annotations

import ctypes
import random
from typing import Dict, Literal, Optional, Sequence, Tuple

import torch
from torch import Tensor

from ._abi import check, lib, ptr
from .knn import knn_sk
from .ops import _f32c, _req_cuda, _stream

N_POINTS_IN_RANGE = 21  # dn_model.py:1780
RANGE_SIZE = 3          # dn_model.py:1779


def level_crossings(points: Tensor, cam_pos: Sequence[float], closest_gaussians: Tensor, means: Tensor, scales: Tensor,
                    quats: Tensor, opacities: Tensor, surface_levels: Sequence[float], return_densities: bool = False):
    """points [P,3], closest_gaussians [P,K] int64 (into the model's Gaussians), the model's raw parameters ->
    (t [L,P] float32 ray parameter of the first crossing of each level, valid [L,P] bool, std [P],
    densities [P,21] or None)."""
    _req_cuda(points, closest_gaussians, means, scales, quats, opacities)
    assert closest_gaussians.dtype == torch.int64 and closest_gaussians.dim() == 2
    P, K = closest_gaussians.shape
    assert points.shape == (P, 3), points.shape
    L = len(surface_levels)
    dev = points.device
    t = torch.empty((L, P), dtype=torch.float32, device=dev)
    valid = torch.empty((L, P), dtype=torch.uint8, device=dev)
    std = torch.empty((P,), dtype=torch.float32, device=dev)
    dens = torch.empty((P, N_POINTS_IN_RANGE), dtype=torch.float32, device=dev) if return_densities else None
    lin = torch.linspace(-RANGE_SIZE, RANGE_SIZE, N_POINTS_IN_RANGE)  # dn_model.py:1784-1786, the reference's values
    c_lin = (ctypes.c_float * N_POINTS_IN_RANGE)(*[float(v) for v in lin])
    c_cam = (ctypes.c_float * 3)(*[float(v) for v in cam_pos])
    c_lev = (ctypes.c_float * L)(*[float(v) for v in surface_levels])
    check(lib.fsb_level_crossings(P, ptr(_f32c(points.detach())), ctypes.addressof(c_cam), K,
                                  ptr(closest_gaussians.contiguous()), ptr(_f32c(means.detach())),
                                  ptr(_f32c(scales.detach())), ptr(_f32c(quats.detach())), ptr(_f32c(opacities.detach())),
                                  ctypes.addressof(c_lin), L, ctypes.addressof(c_lev), ptr(t), ptr(valid), ptr(dens),
                                  ptr(std), _stream()), "fsb_level_crossings")
    return t, valid.bool(), std, dens


def backproject_depth(depth: Tensor, fx: float, fy: float, cx: float, cy: float, W: int, H: int, c2w: Tensor) -> Tensor:
    """get_means3d_backproj (dn_splatter/utils/camera_utils.py:92-144): pixel centres (u + 0.5, v + 0.5), row-major,
    `(u - cx) * d / fx`, then `p @ inv(R) + t` with the OpenCV-convention c2w [3,4]."""
    d = depth.reshape(-1).float()
    v, u = torch.meshgrid(torch.arange(H, device=depth.device), torch.arange(W, device=depth.device), indexing="ij")
    u = (u.reshape(-1) + 0.5).float()
    v = (v.reshape(-1) + 0.5).float()
    pts = torch.stack([(u - cx) * d / fx, (v - cy) * d / fy, d], dim=-1)
    return pts @ torch.linalg.inv(c2w[..., :3, :3].float()) + c2w[..., :3, 3].float()


@torch.no_grad()
def compute_level_surface_points(model, camera, num_samples: int, mask: Optional[Tensor] = None,
                                 surface_levels: Tuple[float, ...] = (0.1, 0.3, 0.5),
                                 return_normal: Literal["analytical", "closest_gaussian", "average"] = "closest_gaussian",
                                 ) -> Dict[float, Dict[str, Tensor]]:
    """dn_model.py:1705-1946 with `model` in the place of `self` (same arguments, same result dict:
    {level: {"points", "normals", "colors"}} of at most `num_samples` random rows each)."""
    if return_normal != "closest_gaussian":
        raise NotImplementedError("level_set: return_normal='closest_gaussian' (the reference's default) only")
    c2w = camera.camera_to_worlds.squeeze(0)
    c2w = c2w @ torch.diag(torch.tensor([1, -1, -1, 1], device=c2w.device, dtype=c2w.dtype))  # :1728-1731
    outputs = model.get_outputs(camera=camera)
    depth, rgb = outputs["depth"], outputs["rgb"]
    W, H = int(camera.width.item()), int(camera.height.item())
    points = backproject_depth(depth, camera.fx.item(), camera.fy.item(), camera.cx.item(), camera.cy.item(), W, H, c2w)
    points = points.view(H, W, -1)
    colors = rgb.reshape(-1, 3).view(H, W, 3)
    if mask is not None:  # :1752-1755
        mask = mask.to(points.device)
        points = points * mask
        depth = depth * mask
    no_depth_mask = (depth <= 0.0)[..., 0]
    points = points[~no_depth_mask].contiguous()
    colors = colors[~no_depth_mask]
    k = int(model.config.knn_to_track)
    closest = knn_sk(model.means.data, points, k=k)  # :1761-1763
    cam_pos = camera.camera_to_worlds.detach()[..., :3, 3].reshape(3)
    t, valid, _, _ = level_crossings(points, cam_pos.tolist(), closest, model.means, model.scales, model.quats,
                                     model.opacities, surface_levels)
    camera_to_samples = torch.nn.functional.normalize(points - cam_pos, dim=-1)
    all_outputs = {}
    for li, level in enumerate(surface_levels):
        keep = valid[li]
        intersection_points = points[keep] + t[li][keep][:, None] * camera_to_samples[keep]  # :1876-1879
        intersection_colors = colors[keep]
        intersection_normals = model.normals[closest[keep][..., 0]]  # :1916-1920
        n = intersection_points.shape[0]
        indices = random.sample(range(n), num_samples if num_samples < n else n)  # :1926-1931
        idx = torch.tensor(indices, device=points.device, dtype=torch.long)
        all_outputs[level] = {"points": intersection_points[idx], "normals": intersection_normals[idx],
                              "colors": intersection_colors[idx]}
    return all_outputs
