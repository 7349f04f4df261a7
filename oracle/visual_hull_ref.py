"""CPU oracle for voxel carving: a numpy restatement of the reference's utils/VisualHull.py.

TEST INFRASTRUCTURE ONLY (tests/, smoke(), bench.py's CPU-baseline legs); never imported by fusionsense_b200/.

Pinned: tests/golden/visual_hull_201.npz holds outputs of the UNMODIFIED reference function
/root/reference/utils/VisualHull.py:87-200 run in this container on a synthetic capture
(generator: oracle/make_golden_visual_hull.py); tests/test_oracle_visual_hull.py checks this restatement
against them bit for bit (occupied coordinates, vote maximum, threshold).

Each function cites the reference lines it follows.  Differences are structural only: the three pure-Python
voxel loops become broadcasts, and grid size / extent are parameters (the reference hard-codes a +-0.5 m cube
with 5 mm voxels, VisualHull.py:135-145) so BASELINE.json's 512^3 configuration can be expressed.
"""
from __future__ import annotations

import json
import os
from dataclasses import dataclass
from pathlib import Path
from typing import List, Optional, Sequence, Tuple

import numpy as np


@dataclass
class HullCameras:
    mats: np.ndarray  # [M,3,4] float64   K @ [R|t]            (VisualHull.py:107-116)
    camera_center: np.ndarray  # [3] float32  mean of c2w translations (VisualHull.py:112-118)
    names: List[str]  # image stems, order of transforms.json frames filtered by train_filenames


def cameras_from_transforms(path: str, transformsfile: str = "transforms.json") -> HullCameras:
    """utils/readCam.py:18-55 (pose part only) + utils/VisualHull.py:92-118."""
    with open(os.path.join(path, transformsfile)) as f:
        contents = json.load(f)
    FX, FY = np.float32(contents["fl_x"]), np.float32(contents["fl_y"])
    CX, CY = np.float32(contents["cx"]), np.float32(contents["cy"])
    K = np.eye(3, dtype=np.float32)
    K[0, 0], K[1, 1], K[0, 2], K[1, 2] = FX, FY, CX, CY
    mats, poses, names = [], [], []
    w2c32 = np.eye(4, dtype=np.float32)
    for frame in contents["frames"]:
        cam_name = os.path.join(frame["file_path"])
        if cam_name not in contents["train_filenames"]:  # readCam.py:27-28
            continue
        c2w = np.array(frame["transform_matrix"])
        w2c = np.linalg.inv(c2w)  # readCam.py:36
        R_stored = np.transpose(w2c[:3, :3])  # readCam.py:37
        T = w2c[:3, 3]
        R = R_stored.T  # VisualHull.py:109
        t = T.reshape(3, 1)
        w2c32[:3, :3] = R  # VisualHull.py:112-113 (float32 storage)
        w2c32[:3, 3] = T
        c2w32 = np.linalg.inv(w2c32)
        poses.append(c2w32[:3, 3])
        mats.append(np.matmul(K, np.concatenate([R, t], axis=1)))  # VisualHull.py:116
        names.append(Path(cam_name).stem)
    camera_center = np.mean(poses, axis=0)  # VisualHull.py:118
    return HullCameras(np.stack(mats).astype(np.float64), camera_center, names)


def load_masks(path: str, names: Sequence[str]) -> np.ndarray:
    """VisualHull.py:121-133 -> uint8 [M,H,W] (the reference divides by 255 into float64 [H,W,M])."""
    import cv2

    out = []
    for name in names:
        m = cv2.imread(os.path.join(path, "masks", f"{name}.png"), cv2.IMREAD_UNCHANGED)
        if m.ndim == 3:
            m = m[:, :, 0]
        out.append(m)
    return np.stack(out)


def grid_axes(camera_center: np.ndarray, half_extent: float = 0.5, voxel_size: float = 0.005,
              n_per_axis: Optional[int] = None) -> Tuple[np.ndarray, np.ndarray, np.ndarray]:
    """VisualHull.py:135-147 + InitializeVoxels :15-57 -> (xs, ys, zs) as float64, zs in LOOP order (descending).

    The limits are computed exactly as the reference does (camera_center is float32; what `float32 - 0.5`
    promotes to depends on the running numpy), so the axes are whatever np.linspace yields
    here; they are widened to float64 the same way `voxel[l] = [x, y, z, 1]` does (VisualHull.py:54).
    `n_per_axis` overrides the count (for the 512^3 configuration); default int(|hi-lo|/voxel)+1.
    """
    lims = []
    for a in range(3):
        lo = camera_center[a] - half_extent
        hi = camera_center[a] + half_extent
        lims.append((lo, hi))
    counts = []
    for lo, hi in lims:
        n = np.abs(hi - lo) / voxel_size
        counts.append(int(np.array(n).astype(int)) + 1 if n_per_axis is None else int(n_per_axis))
    (sx, ex), (sy, ey), (sz, ez) = lims
    xs = np.linspace(sx, ex, counts[0])
    ys = np.linspace(sy, ey, counts[1])
    zs = np.linspace(ez, sz, counts[2])  # z runs from max to min in the outer loop (:51)
    return xs.astype(np.float64), ys.astype(np.float64), zs.astype(np.float64)


def project_votes(mats: np.ndarray, masks_u8: np.ndarray, xs: np.ndarray, ys: np.ndarray, zs: np.ndarray,
                  z_chunk: int = 8) -> np.ndarray:
    """VisualHull.py:154-172 -> votes float64 [nz*nx*ny] in the reference's voxel order (z outer, x, y inner)."""
    M, H, W = masks_u8.shape
    nx, ny, nz = len(xs), len(ys), len(zs)
    votes = np.zeros(nz * nx * ny, dtype=np.float64)
    imgs = [np.array(masks_u8[i] / 255) for i in range(M)]  # :133 (float64)
    X = np.repeat(xs, ny)
    Y = np.tile(ys, nx)
    for z0 in range(0, nz, z_chunk):
        zc = zs[z0:z0 + z_chunk]
        n = len(zc) * nx * ny
        P = np.ones((n, 4))  # rows [x, y, z, 1] like `voxel` (:24,54), transposed view below like :149
        P[:, 0] = np.tile(X, len(zc))
        P[:, 1] = np.tile(Y, len(zc))
        P[:, 2] = np.repeat(zc, nx * ny)
        pts = P.T
        acc = votes[z0 * nx * ny: z0 * nx * ny + n]
        for i in range(M):
            p2 = np.matmul(mats[i], pts)  # :158
            with np.errstate(divide="ignore", invalid="ignore"):
                p2 = np.floor(p2 / p2[2, :] + 1e-6).astype(np.int32)  # :159
            p2[np.where(p2 < 0)] = 0  # :160
            ind = np.where(p2[1, :] >= H)  # :163
            p2[:, ind] = 0
            ind = np.where(p2[0, :] >= W)  # :165
            p2[:, ind] = 0
            acc += imgs[i].T[p2.T[:, 0], p2.T[:, 1]]  # :170
    return votes


def threshold(votes: np.ndarray, error: float = 5) -> Tuple[float, float]:
    """VisualHull.py:174-176 -> (maxv, iso_value)."""
    maxv = np.max(votes)
    iso_value = maxv - np.round(((maxv) / 100) * error) - 0.5
    return float(maxv), float(iso_value)


def occupied_points(votes: np.ndarray, iso_value: float, xs, ys, zs) -> np.ndarray:
    """VisualHull.py:185-191 -> float64 [n_occ,3] xyz in voxel order."""
    nx, ny = len(xs), len(ys)
    idx = np.nonzero(votes > iso_value)[0]
    iz, rem = np.divmod(idx, nx * ny)
    ix, iy = np.divmod(rem, ny)
    return np.stack([xs[ix], ys[iy], zs[iz]], axis=1)


def visual_hull(path: str, error: float = 5, n_per_axis: Optional[int] = None, half_extent: float = 0.5,
                voxel_size: float = 0.005):
    """Whole pipeline of VisualHull.py:87-191 minus file output/plotting. -> (points [n_occ,3] f64, maxv, iso)."""
    cams = cameras_from_transforms(path)
    masks = load_masks(path, cams.names)
    xs, ys, zs = grid_axes(cams.camera_center, half_extent, voxel_size, n_per_axis)
    votes = project_votes(cams.mats, masks, xs, ys, zs)
    maxv, iso = threshold(votes, error)
    return occupied_points(votes, iso, xs, ys, zs), maxv, iso
