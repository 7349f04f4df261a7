"""Seeded synthetic scenes shaped like FusionSense captures.

Used by tests, bench.py and smoke(); there is no network for real datasets.  Everything is generated on
the CPU with a fixed seed and then moved to the requested device, so the oracle and the kernels see
bit-identical inputs.
"""
from __future__ import annotations

import math
from dataclasses import dataclass
from typing import Optional

import torch
from torch import Tensor

SEED_BASE = 20241011
SH_C0 = 0.28209479177387814


def rgb_to_sh(rgb: Tensor) -> Tensor:
    """nerfstudio `RGB2SH` (used at /root/reference/dn_splatter/dn_model.py:211,1193)."""
    return (rgb - 0.5) / SH_C0


@dataclass
class Scene:
    means: Tensor  # [N,3]
    scales: Tensor  # [N,3] log-scales (DN-Splatter parameterisation, dn_model.py:268-269)
    quats: Tensor  # [N,4] wxyz, un-normalised
    opacities: Tensor  # [N,1] logits
    features_dc: Tensor  # [N,3]
    features_rest: Tensor  # [N,15,3]
    viewmats: Tensor  # [M,4,4] world->camera (OpenCV)
    c2w: Tensor  # [M,4,4]
    Ks: Tensor  # [M,3,3]
    width: int
    height: int

    def to(self, device) -> "Scene":
        kw = {k: (v.to(device) if isinstance(v, Tensor) else v) for k, v in self.__dict__.items()}
        return Scene(**kw)

    @property
    def N(self) -> int:
        return self.means.shape[0]


def look_at_cameras(n_views: int, radius: float, elevation_deg: float = 30.0, dtype=torch.float32):
    """Ring of OpenCV cameras (x right, y down, z forward) looking at the origin. -> c2w [M,4,4], w2c [M,4,4]."""
    c2ws = []
    el = math.radians(elevation_deg)
    for i in range(n_views):
        az = 2 * math.pi * i / n_views
        pos = torch.tensor([radius * math.cos(el) * math.cos(az), radius * math.cos(el) * math.sin(az),
                            radius * math.sin(el)], dtype=torch.float64)
        fwd = -pos / pos.norm()
        up = torch.tensor([0.0, 0.0, 1.0], dtype=torch.float64)
        right = torch.linalg.cross(fwd, up)
        right = right / right.norm()
        down = torch.linalg.cross(fwd, right)
        c2w = torch.eye(4, dtype=torch.float64)
        c2w[:3, 0], c2w[:3, 1], c2w[:3, 2], c2w[:3, 3] = right, down, fwd, pos
        c2ws.append(c2w)
    c2w = torch.stack(c2ws)
    w2c = torch.linalg.inv(c2w)
    return c2w.to(dtype), w2c.to(dtype)


def bunny_points(n: int, gen: torch.Generator, scale: float = 1.0) -> Tensor:
    """Points on a procedural 'bunny': the union of three ellipsoid shells (body, head, ears blob)."""
    parts = [((0.0, 0.0, 0.0), (0.45, 0.32, 0.30), 0.55), ((0.42, 0.0, 0.28), (0.20, 0.17, 0.17), 0.25),
             ((0.50, 0.0, 0.55), (0.06, 0.12, 0.22), 0.20)]
    out = []
    for k, (ctr, rad, frac) in enumerate(parts):
        m = n - sum(o.shape[0] for o in out) if k == len(parts) - 1 else int(n * frac)
        d = torch.randn(m, 3, generator=gen)
        d = d / d.norm(dim=-1, keepdim=True)
        out.append(d * torch.tensor(rad) + torch.tensor(ctr))
    return torch.cat(out) * scale


def make_scene(n_gaussians: int, width: int, height: int, n_views: int = 1, cfg_id: int = 1, kind: str = "random",
               fx: Optional[float] = None, cam_radius: Optional[float] = None, sh_degree: int = 3) -> Scene:
    """kind='random': means ~ U([-1,1]^3), cameras at radius 2.5.  kind='bunny': RealSense-shaped object scene
    (object ~0.12 m, cameras at 0.4 m, 20 % background shell) as in BASELINE.json configs[1]."""
    g = torch.Generator().manual_seed(SEED_BASE + cfg_id)
    N = n_gaussians
    if kind == "random":
        means = torch.rand(N, 3, generator=g) * 2 - 1
        base = 0.01 * (1e6 / N) ** (1 / 3)
        cam_radius = 2.5 if cam_radius is None else cam_radius
    elif kind == "bunny":
        # 30 % of the Gaussians on the object, 70 % on the surroundings (table plane + room shell), like a
        # FusionSense capture where most of the ~300k Gaussians model the background.
        n_obj = (3 * N) // 10
        n_floor = (N - n_obj) // 2
        n_wall = N - n_obj - n_floor
        obj = bunny_points(n_obj, g, scale=0.12)
        floor = torch.stack([(torch.rand(n_floor, generator=g) * 2 - 1) * 0.9,
                             (torch.rand(n_floor, generator=g) * 2 - 1) * 0.9,
                             torch.full((n_floor,), -0.045)], dim=-1)
        d = torch.randn(n_wall, 3, generator=g)
        d[:, 2] = d[:, 2].abs()
        wall = d / d.norm(dim=-1, keepdim=True) * 0.95
        means = torch.cat([obj, floor, wall])
        base = torch.cat([torch.full((n_obj,), 0.0008), torch.full((N - n_obj,), 0.004)]) * (3e5 / N) ** 0.5
        cam_radius = 0.4 if cam_radius is None else cam_radius
    else:
        raise ValueError(kind)
    base_t = base if isinstance(base, Tensor) else torch.full((N,), float(base))
    log_s = torch.log(base_t)[:, None] + 0.3 * torch.randn(N, 3, generator=g)
    log_s[:, 2] += math.log(0.1)  # surfel-like third axis
    quats = torch.randn(N, 4, generator=g)
    opac = 1.5 * torch.randn(N, 1, generator=g)
    dc = rgb_to_sh(torch.rand(N, 3, generator=g))
    K = (sh_degree + 1) ** 2
    rest = 0.05 * torch.randn(N, max(K - 1, 0), 3, generator=g)
    c2w, w2c = look_at_cameras(n_views, cam_radius)
    fx = fx if fx is not None else {640: 600.0, 1920: 1400.0, 3840: 2800.0}.get(width, 0.9375 * width)
    Ks = torch.tensor([[fx, 0, width / 2], [0, fx, height / 2], [0, 0, 1]], dtype=torch.float32)[None].repeat(
        n_views, 1, 1)
    return Scene(means.float(), log_s.float(), quats.float(), opac.float(), dc.float(), rest.float(), w2c, c2w, Ks,
                 width, height)


# ---------------------------------------------------------------------------------------------
# synthetic capture on disk in the FusionSense layout — VisualHull input
# ---------------------------------------------------------------------------------------------
BUNNY_PARTS = (((0.0, 0.0, 0.0), (0.45, 0.32, 0.30)), ((0.42, 0.0, 0.28), (0.20, 0.17, 0.17)),
               ((0.50, 0.0, 0.55), (0.06, 0.12, 0.22)))


def bunny_silhouettes(c2w, fx: float, fy: float, cx: float, cy: float, width: int, height: int, scale: float = 0.12,
                      center=(0.0, 0.0, 0.0)):
    """uint8 0/255 masks [M,H,W]: exact ray / ellipsoid-union intersection of the procedural bunny solid."""
    import numpy as np

    c2w = np.asarray(c2w, dtype=np.float64)
    js, is_ = np.meshgrid(np.arange(width) + 0.5, np.arange(height) + 0.5)
    d_cam = np.stack([(js - cx) / fx, (is_ - cy) / fy, np.ones_like(js)], -1)  # OpenCV camera rays
    masks = []
    for m in range(c2w.shape[0]):
        R, o = c2w[m, :3, :3], c2w[m, :3, 3]
        d = d_cam @ R.T
        hit = np.zeros((height, width), bool)
        for ctr, rad in BUNNY_PARTS:
            ctr = np.asarray(ctr) * scale + np.asarray(center)
            rad = np.asarray(rad) * scale
            oo = (o - ctr) / rad
            dd = d / rad
            a = (dd * dd).sum(-1)
            b = 2 * (dd * oo).sum(-1)
            c = (oo * oo).sum() - 1.0
            disc = b * b - 4 * a * c
            t = (-b - np.sqrt(np.maximum(disc, 0))) / (2 * a)
            hit |= (disc >= 0) & (t > 0)
        masks.append(hit.astype(np.uint8) * 255)
    return np.stack(masks)


def write_capture(path, n_views: int = 9, width: int = 640, height: int = 480, fx: float = 600.0,
                  cam_radius: float = 0.4, object_center=(0.02, -0.01, 0.03), masks=None, c2w=None):
    """Write transforms.json + images/ + masks/ the way utils/VisualHull.py and utils/readCam.py read them.

    Returns (c2w [M,4,4] float64 numpy, masks uint8 [M,H,W]).
    """
    import json
    import os

    import cv2
    import numpy as np

    os.makedirs(os.path.join(path, "images"), exist_ok=True)
    os.makedirs(os.path.join(path, "masks"), exist_ok=True)
    if c2w is None:
        c2w_t, _ = look_at_cameras(n_views, cam_radius, elevation_deg=30.0, dtype=torch.float64)
        c2w = c2w_t.numpy().copy()
        c2w[:, :3, 3] += np.asarray(object_center)
    cx, cy = width / 2.0, height / 2.0
    if masks is None:
        masks = bunny_silhouettes(c2w, fx, fx, cx, cy, width, height, center=object_center)
    frames, names = [], []
    for i in range(c2w.shape[0]):
        fp = f"images/rgb_{i}.png"
        names.append(fp)
        frames.append({"file_path": fp, "transform_matrix": c2w[i].tolist()})
        rgb = np.repeat(masks[i][..., None], 3, axis=-1)
        cv2.imwrite(os.path.join(path, fp), rgb)
        cv2.imwrite(os.path.join(path, "masks", f"rgb_{i}.png"), masks[i])
    meta = {"fl_x": fx, "fl_y": fx, "cx": cx, "cy": cy, "w": width, "h": height, "frames": frames,
            "train_filenames": names, "val_filenames": [], "test_filenames": []}
    with open(os.path.join(path, "transforms.json"), "w") as f:
        json.dump(meta, f)
    return c2w, masks
