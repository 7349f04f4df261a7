"""TEST INFRASTRUCTURE — CPU restatement of the reference's nearest-neighbour search and density field.

* `knn_sk_ref(x, y, k)`  — dn_splatter/utils/knn.py:29-43: the k + 1 nearest rows of x for every row of y (Euclidean,
  float64 arithmetic on the float32 coordinates like the KD-tree sklearn's `algorithm="auto"` picks for 3-D data),
  nearest first, FIRST COLUMN DROPPED.  Brute force in numpy; ties are ordered by index (sklearn's tie order is an
  implementation detail of its heap — the tests compare distances wherever two candidates tie).
* `get_density_ref(...)` — dn_splatter/dn_model.py:1596-1634 + scale_rot_to_inv_cov3d (:2141-2150) in plain torch.

Pinned: tests/golden/knn_sk.npz is produced by running the reference's own knn.py / dn_model.py here
(oracle/make_golden_knn.py).  Only tests/ and tools/ may import this module.
"""
from __future__ import annotations

import numpy as np
import torch


def knn_full_ref(x: np.ndarray, y: np.ndarray, k: int, block: int = 512):
    """The k nearest rows of x for every row of y: (indices [Ny,k] int64, distances [Ny,k] float64), ordered by
    (distance, index)."""
    x64 = np.asarray(x, dtype=np.float32).astype(np.float64)
    y64 = np.asarray(y, dtype=np.float32).astype(np.float64)
    idx = np.empty((len(y64), k), dtype=np.int64)
    dist = np.empty((len(y64), k), dtype=np.float64)
    ar = np.arange(len(x64))
    for s in range(0, len(y64), block):
        q = y64[s:s + block]
        d2 = ((q[:, None, :] - x64[None, :, :]) ** 2).sum(-1)  # exact differences and squares, like the kernel
        order = np.lexsort((np.broadcast_to(ar, d2.shape), d2), axis=-1)[:, :k]
        idx[s:s + block] = order
        dist[s:s + block] = np.sqrt(np.take_along_axis(d2, order, axis=-1))
    return idx, dist


def knn_sk_ref(x, y, k: int) -> np.ndarray:
    """knn.py:29-43: indices [Ny, k] of neighbours 2 .. k + 1."""
    return knn_full_ref(np.asarray(x), np.asarray(y), k + 1)[0][:, 1:]


def quat_to_rotmat_ref(quat: torch.Tensor) -> torch.Tensor:
    """gsplat 0.1.x `quat_to_rotmat` (wxyz, normalised) — what dn_model.py:286 imports."""
    q = torch.nn.functional.normalize(quat, dim=-1)
    w, x, y, z = torch.unbind(q, dim=-1)
    return torch.stack([
        1 - 2 * (y ** 2 + z ** 2), 2 * (x * y - w * z), 2 * (x * z + w * y),
        2 * (x * y + w * z), 1 - 2 * (x ** 2 + z ** 2), 2 * (y * z - w * x),
        2 * (x * z - w * y), 2 * (y * z + w * x), 1 - 2 * (x ** 2 + y ** 2),
    ], dim=-1).reshape(quat.shape[:-1] + (3, 3))


def get_density_ref(samples, closest, means, log_scales, quats, opacities):
    """dn_model.py:1596-1634 with explicit parameters instead of `self`."""
    centers = means[closest]
    scale = 1.0 / torch.exp(log_scales[closest]).clamp(min=1e-3)
    M = quat_to_rotmat_ref(quats[closest]) * scale[..., None, :]
    op = torch.sigmoid(opacities[closest])
    dist = samples[:, None, :] - centers
    man = M.transpose(-1, -2) @ dist[..., None]
    maha = (man[..., 0] * man[..., 0]).sum(dim=-1).clamp(min=0.0, max=1e8)
    dens = (op[..., 0] * torch.exp(-0.5 * maha)).sum(dim=-1)
    mask = dens >= 1.0
    dens = dens.clone()
    dens[mask] = dens[mask] / (dens[mask] + 1e-5)
    return dens.clamp(min=1e-4)
