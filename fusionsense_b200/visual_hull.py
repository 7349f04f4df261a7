"""B200 visual-hull carving behind the reference's interface.

`VisualHull(path, output_path, error=5)` has the name, arguments and on-disk result (`foreground_pcd.ply`)
of /root/reference/utils/VisualHull.py:87-200, so `scripts/train.py:284` can call it unchanged; the carving
itself (projection / votes / threshold / occupied-voxel extraction, :149-191) runs in libfsb200's
`fsb_vh_*` kernels.  Camera / mask loading and the grid limits are host numpy that follows the reference
line by line, because their dtype promotion is numpy-version dependent and must be
whatever the running numpy does.  No CPU fallback for the carving.

Multi-GPU: `carve(..., rank, world_size)` shards the outer (z, descending) voxel axis into contiguous slabs,
one per rank; `carve_distributed` adds the MAX all-reduce of the vote maximum and the
all-gather of the occupied points in rank order (= reference order).
"""
from __future__ import annotations

import json
import os
from pathlib import Path
from typing import List, Optional, Tuple

import numpy as np
import torch

from ._abi import check, lib, ptr
from .ops import _stream, isect_scan


# ---------------------------------------------------------------------------------------------
# host-side input preparation (follows the reference; cheap, runs once)
# ---------------------------------------------------------------------------------------------
def read_hull_cameras(path: str, transformsfile: str = "transforms.json"):
    """-> (mats [M,3,4] float64 = K @ [R|t], camera_center [3] float32, image stems).

    utils/readCam.py:18-55 restricted to what VisualHull consumes (R, T, image_name of the frames listed in
    train_filenames) and utils/VisualHull.py:92-118.
    """
    with open(os.path.join(path, transformsfile)) as f:
        contents = json.load(f)
    K = np.eye(3, dtype=np.float32)
    K[0, 0] = np.float32(contents["fl_x"])
    K[1, 1] = np.float32(contents["fl_y"])
    K[0, 2] = np.float32(contents["cx"])
    K[1, 2] = np.float32(contents["cy"])
    train = contents["train_filenames"]
    w2c_f32 = np.eye(4, dtype=np.float32)
    mats, centres, names = [], [], []
    for frame in contents["frames"]:
        name = os.path.join(frame["file_path"])
        if name not in train:
            continue
        w2c = np.linalg.inv(np.array(frame["transform_matrix"]))
        R, T = w2c[:3, :3], w2c[:3, 3]  # the reference stores R transposed (readCam.py:37) and undoes it (:109)
        w2c_f32[:3, :3] = R
        w2c_f32[:3, 3] = T
        centres.append(np.linalg.inv(w2c_f32)[:3, 3])
        mats.append(np.matmul(K, np.concatenate([R, T.reshape(3, 1)], axis=1)))
        names.append(Path(name).stem)
    if not mats:
        raise ValueError(f"no frame of {transformsfile} is listed in train_filenames")
    return np.stack(mats).astype(np.float64), np.mean(centres, axis=0), names


def read_masks(path: str, names: List[str]) -> np.ndarray:
    """masks/<stem>.png -> uint8 [M,H,W] (first channel), utils/VisualHull.py:121-133."""
    import cv2

    out = []
    for n in names:
        fn = os.path.join(path, "masks", f"{n}.png")
        img = cv2.imread(fn, cv2.IMREAD_UNCHANGED)
        if img is None:
            raise FileNotFoundError(fn)
        out.append(img[:, :, 0] if img.ndim == 3 else img)
    return np.ascontiguousarray(np.stack(out), dtype=np.uint8)


def hull_grid(camera_center, half_extent: float = 0.5, voxel_size: float = 0.005, n_per_axis: Optional[int] = None):
    """Axis tables (xs, ys, zs) as float64, zs descending — utils/VisualHull.py:135-147 and :15-57.

    `n_per_axis` replaces the reference's int(|hi-lo|/voxel)+1 (201 for its hard-coded +-0.5 m / 5 mm grid).
    """
    axes = []
    for a in range(3):
        lo, hi = camera_center[a] - half_extent, camera_center[a] + half_extent
        n = int(np.array(np.abs(hi - lo) / voxel_size).astype(int)) + 1 if n_per_axis is None else int(n_per_axis)
        axes.append((lo, hi, n))
    xs = np.linspace(axes[0][0], axes[0][1], axes[0][2])
    ys = np.linspace(axes[1][0], axes[1][1], axes[1][2])
    zs = np.linspace(axes[2][1], axes[2][0], axes[2][2])
    return xs.astype(np.float64), ys.astype(np.float64), zs.astype(np.float64)


def slab_bounds(nz: int, rank: int, world_size: int) -> Tuple[int, int]:
    """Contiguous slab [z0, z1) of the outer voxel axis owned by `rank` (sizes differ by at most one)."""
    base, rem = divmod(nz, world_size)
    z0 = rank * base + min(rank, rem)
    return z0, z0 + base + (1 if rank < rem else 0)


def iso_value(maxv: float, error: float) -> float:
    """utils/VisualHull.py:174-176 (np.round = round-half-to-even)."""
    maxv = np.float64(maxv)
    return float(maxv - np.round(((maxv) / 100) * error) - 0.5)


# ---------------------------------------------------------------------------------------------
# device side
# ---------------------------------------------------------------------------------------------
class HullCarver:
    """Holds the device-resident inputs (masks, axis tables, lookup table) of one carving job."""

    def __init__(self, mats: np.ndarray, masks_u8: np.ndarray, xs, ys, zs, device="cuda", rank: int = 0,
                 world_size: int = 1):
        if not torch.cuda.is_available():
            raise RuntimeError("fusionsense_b200.visual_hull needs a CUDA device (no CPU fallback)")
        M, H, W = masks_u8.shape
        if M > lib.fsb_vh_max_views():
            raise ValueError(f"at most {lib.fsb_vh_max_views()} views per launch")
        self.device = torch.device(device)
        self.mats = np.ascontiguousarray(mats.reshape(M, 12), dtype=np.float64)
        self.M, self.H, self.W = M, H, W
        z0, z1 = slab_bounds(len(zs), rank, world_size)
        self.z0, self.z1 = z0, z1
        self.nx, self.ny, self.nz = len(xs), len(ys), z1 - z0
        self.masks = torch.from_numpy(masks_u8).to(self.device)
        self.xs = torch.from_numpy(np.ascontiguousarray(xs)).to(self.device)
        self.ys = torch.from_numpy(np.ascontiguousarray(ys)).to(self.device)
        self.zs = torch.from_numpy(np.ascontiguousarray(zs[z0:z1])).to(self.device)
        # mask value -> vote, computed by numpy exactly like `mask_img/255` (VisualHull.py:133)
        self.lut = torch.from_numpy(np.arange(256, dtype=np.uint8) / 255).to(self.device)
        self.V = self.nz * self.nx * self.ny
        self.votes = None
        # binary masks (what Grounded-SAM-2 writes): a vote is an exact integer count, stored in one byte per voxel
        self.votes_u8 = bool(M <= 255 and np.isin(masks_u8, (0, 255)).all())

    def vote(self) -> float:
        """Launch the vote kernel over this rank's slab; returns the slab's vote maximum (one 8-byte D2H)."""
        self.votes = torch.empty((self.V,), dtype=torch.uint8 if self.votes_u8 else torch.float64, device=self.device)
        max_bits = torch.zeros((1,), dtype=torch.int64, device=self.device)
        check(lib.fsb_vh_votes(self.M, self.H, self.W, ptr(self.masks), self.mats.ctypes.data, ptr(self.lut),
                               ptr(self.xs), self.nx, ptr(self.ys), self.ny, ptr(self.zs), self.nz, ptr(self.votes),
                               int(self.votes_u8), ptr(max_bits), _stream()), "fsb_vh_votes")
        return float(max_bits.view(torch.float64).item())

    def extract(self, iso: float, want_indices: bool = False):
        """Occupied voxels (votes > iso) of this slab in voxel order -> points [n_occ,3] float64 (device)."""
        if self.votes is None:
            raise RuntimeError("call vote() first")
        nb = (self.V + lib.fsb_vh_count_block() - 1) // lib.fsb_vh_count_block()
        counts = torch.empty((max(nb, 1),), dtype=torch.int32, device=self.device)
        if self.V == 0:
            counts.zero_()
        check(lib.fsb_vh_count(self.V, ptr(self.votes), int(self.votes_u8), iso, ptr(counts), _stream()), "fsb_vh_count")
        offsets, n_occ = isect_scan(counts)
        points = torch.empty((n_occ, 3), dtype=torch.float64, device=self.device)
        idx = torch.empty((n_occ,), dtype=torch.int64, device=self.device) if want_indices else None
        if n_occ:
            check(lib.fsb_vh_compact(self.V, ptr(self.votes), int(self.votes_u8), iso, ptr(offsets), ptr(self.xs), self.nx, ptr(self.ys),
                                     self.ny, ptr(self.zs), ptr(points), ptr(idx), _stream()), "fsb_vh_compact")
        if want_indices:
            return points, idx + self.z0 * self.nx * self.ny
        return points


def carve(mats, masks_u8, xs, ys, zs, error: float = 5, device="cuda"):
    """Single-GPU carving -> (points [n_occ,3] float64 on device, maxv, iso)."""
    c = HullCarver(mats, masks_u8, xs, ys, zs, device=device)
    maxv = c.vote()
    iso = iso_value(maxv, error)
    return c.extract(iso), maxv, iso


def gather_slabs(local_points: torch.Tensor, group=None) -> torch.Tensor:
    """All-gather variable-length [n_r,3] point blocks and concatenate them in rank order."""
    import torch.distributed as dist

    world = dist.get_world_size(group)
    n_local = torch.tensor([local_points.shape[0]], dtype=torch.int64, device=local_points.device)
    counts = [torch.zeros_like(n_local) for _ in range(world)]
    dist.all_gather(counts, n_local, group=group)
    counts = [int(c.item()) for c in counts]
    n_max = max(counts) if counts else 0
    padded = torch.zeros((n_max, 3), dtype=local_points.dtype, device=local_points.device)
    padded[: local_points.shape[0]] = local_points
    parts = [torch.empty_like(padded) for _ in range(world)]
    dist.all_gather(parts, padded, group=group)
    return torch.cat([p[:c] for p, c in zip(parts, counts)], dim=0)


def reduce_max(local_max: float, device, group=None) -> float:
    import torch.distributed as dist

    t = torch.tensor([local_max], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX, group=group)
    return float(t.item())


def carve_distributed(mats, masks_u8, xs, ys, zs, error: float = 5, device="cuda", group=None, gather: bool = True):
    """Voxel-slab sharded carving over the ranks of `group` (torch.distributed must be initialised).

    No data-path collective besides the 8-byte MAX of the vote maximum; the final all-gather only assembles
    the result and can be skipped with gather=False (each rank then keeps its slab's points).
    """
    import torch.distributed as dist

    rank, world = dist.get_rank(group), dist.get_world_size(group)
    c = HullCarver(mats, masks_u8, xs, ys, zs, device=device, rank=rank, world_size=world)
    maxv = reduce_max(c.vote(), c.device, group)
    iso = iso_value(maxv, error)
    pts = c.extract(iso)
    return (gather_slabs(pts, group) if gather else pts), maxv, iso


# ---------------------------------------------------------------------------------------------
# the reference-facing entry point
# ---------------------------------------------------------------------------------------------
def write_ply_points(filename: str, points: np.ndarray) -> None:
    """Binary little-endian PLY with double x/y/z — what open3d's write_point_cloud emits for a
    points-only cloud (utils/VisualHull.py:191-193)."""
    pts = np.ascontiguousarray(points, dtype="<f8")
    header = ("ply\nformat binary_little_endian 1.0\ncomment Created by fusionsense_b200\n"
              f"element vertex {pts.shape[0]}\nproperty double x\nproperty double y\nproperty double z\nend_header\n")
    with open(filename, "wb") as f:
        f.write(header.encode("ascii"))
        f.write(pts.tobytes())


def VisualHull(path, output_path, error=5, n_per_axis: Optional[int] = None, device="cuda"):
    """Drop-in for utils/VisualHull.py::VisualHull: reads <path>/transforms.json and <path>/masks/*.png,
    writes <output_path>/foreground_pcd.ply.  Returns the occupied points as a float64 numpy array [n_occ,3].
    (The reference also saves a matplotlib scatter plot, voxels.png; that cosmetic output is not produced.)
    """
    mats, camera_center, names = read_hull_cameras(str(path))
    for n in names:
        print(n)
    print("camera_center:", camera_center)
    masks = read_masks(str(path), names)
    xs, ys, zs = hull_grid(camera_center, n_per_axis=n_per_axis)
    if torch.distributed.is_available() and torch.distributed.is_initialized() and torch.distributed.get_world_size() > 1:
        pts, maxv, iso = carve_distributed(mats, masks, xs, ys, zs, error=error, device=device)
    else:
        pts, maxv, iso = carve(mats, masks, xs, ys, zs, error=error, device=device)
    print("max number of votes:" + str(maxv))
    print("threshold for marching cube:" + str(iso))
    pts = pts.cpu().numpy()
    os.makedirs(str(output_path), exist_ok=True)
    write_ply_points(f"{output_path}/foreground_pcd.ply", pts)
    return pts
