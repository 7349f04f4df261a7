"""TEST INFRASTRUCTURE — CPU restatement of the seed point cloud path.

* `get_pointcloud_ref` — utils/generate_pcd.py:15-48 with the `.cuda()` calls dropped (torch CPU, fp32), otherwise
  the reference's statements in the reference's order.
* `voxel_down_sample_ref` — open3d 0.18 `PointCloud::VoxelDownSample` restated in numpy: voxel_min_bound =
  min_bound - voxel_size / 2, voxel index = floor((p - voxel_min_bound) / voxel_size) in fp64, per voxel the fp64 sums of
  points and colours in input order divided by the count.  **Parity unpinned**: open3d (pyproject.toml dependency, not
  vendored) is not installed in this image, and the reference has no tests or golden vectors for it.  open3d returns the
  voxels in hash-map order; this restatement (like the CUDA path) returns them in ascending (x, y, z) voxel order.

Only tests/ and tools/ may import this module.
"""
from __future__ import annotations

import numpy as np
import torch


def get_pointcloud_ref(color, depth, w2c, FX, FY, CX, CY, transform_pts=True, mask=None):
    width, height = color.shape[2], color.shape[1]
    x_grid, y_grid = torch.meshgrid(torch.arange(width).float(), torch.arange(height).float(), indexing="xy")
    xx = (x_grid - CX) / FX
    yy = (y_grid - CY) / FY
    xx = xx.reshape(-1)
    yy = yy.reshape(-1)
    depth_z = depth.reshape(-1)
    pts_cam = torch.stack((xx * depth_z, yy * depth_z, depth_z), dim=-1)
    if transform_pts:
        c2w = torch.inverse(w2c)
        R = c2w[:3, :3]
        T = c2w[:3, 3]
        pts = ((R @ pts_cam.T) + T.unsqueeze(1)).T
    else:
        pts = pts_cam
    cols = torch.permute(color, (1, 2, 0)).reshape(-1, 3)
    point_cld = torch.cat((pts, cols), -1)
    mask1 = (depth_z > 0) & (depth_z < 0.5)
    mask2 = (depth_z > 0.5) & (depth_z < 5)
    return point_cld[mask1], point_cld[mask2]


def voxel_down_sample_ref(rows: np.ndarray, voxel_size: float) -> np.ndarray:
    rows = np.asarray(rows, dtype=np.float32).astype(np.float64)  # Vector3dVector(float32 array): exact widening
    if len(rows) == 0:
        return rows
    pts = rows[:, :3]
    voxel_min_bound = pts.min(axis=0) - voxel_size * 0.5
    index = np.floor((pts - voxel_min_bound) / voxel_size).astype(np.int64)
    uniq, inverse = np.unique(index, axis=0, return_inverse=True)  # ascending lexicographic (x, y, z)
    inverse = inverse.reshape(-1)
    acc = np.zeros((len(uniq), rows.shape[1]), dtype=np.float64)
    cnt = np.zeros((len(uniq),), dtype=np.float64)
    for i in range(len(rows)):  # sequential, like AccumulatedPoint::AddPoint in input order
        acc[inverse[i]] += rows[i]
        cnt[inverse[i]] += 1.0
    return acc / cnt[:, None]
