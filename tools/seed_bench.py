"""f3 measurement: seed point cloud of 9 RealSense-shaped views (640x480) — back-projection + per-view voxel down-sample
(0.02 m) through fusionsense_b200.seed_points, CUDA events; beside it the CPU oracle (generate_pcd.py's statements in
torch-CPU + the numpy restatement of open3d's VoxelDownSample) on ONE view.  One JSON line on stdout."""
import json
import os
import sys
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))

import torch


def main():
    from fusionsense_b200.seed_points import merged_background_cloud
    from oracle import seed_points_ref as ref

    H, W, V = 480, 640, 9
    g = torch.Generator().manual_seed(4)
    views = []
    for v in range(V):
        color = torch.rand(3, H, W, generator=g)
        depth = 0.6 + 2.0 * torch.rand(H, W, generator=g)
        depth[torch.rand(H, W, generator=g) < 0.2] = 0.3
        w2c = torch.eye(4)
        w2c[:3, 3] = torch.tensor([0.05 * v, 0.0, 0.1])
        views.append((color, depth, w2c))
    dev_views = [(c.cuda(), d.cuda(), w.cuda()) for c, d, w in views]
    fx = fy = 600.0
    cx, cy = 320.0, 240.0
    merged_background_cloud(dev_views, fx, fy, cx, cy)  # warm-up
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    reps = 5
    for _ in range(reps):
        out = merged_background_cloud(dev_views, fx, fy, cx, cy)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    t0 = time.perf_counter()
    _, back = ref.get_pointcloud_ref(*views[0], fx, fy, cx, cy)
    t_bp = time.perf_counter() - t0
    t0 = time.perf_counter()
    down = ref.voxel_down_sample_ref(back.numpy(), 0.02)
    t_vx = time.perf_counter() - t0
    px = H * W * V
    print(json.dumps({
        "metric": "seed_cloud_pixels_per_s", "workload": f"{V} views {W}x{H}: back-projection + voxel_down_sample(0.02)",
        "value": px / (ms * 1e-3), "unit": "pixel/s", "ms_total": ms, "points_out": int(out.shape[0]),
        "algorithmic_bytes": px * (4 + 12 + 24) + int(out.shape[0]) * 48,
        "cpu_baseline": {"kind": "port", "cores": os.cpu_count(), "sample": "1 of 9 views, scaled",
                         "backproject_s_per_view": t_bp, "voxel_s_per_view": t_vx, "value": H * W / (t_bp + t_vx),
                         "unit": "pixel/s", "points_out_view0": int(down.shape[0])},
    }))


if __name__ == "__main__":
    main()
