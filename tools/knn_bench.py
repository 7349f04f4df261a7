"""f2 measurement: exact 16-NN of a 1M-point cloud against itself (the `recompute_knn` call of dn_model.py:172-196 at
cfg4's Gaussian count) — index build + query through fusionsense_b200.knn.knn_sk, CUDA events; beside it the
reference's knn_sk (sklearn on the host cores) on a bounded sample of the queries.  One JSON line on stdout.

  python tools/knn_bench.py [n=1000000] [k=16] [kind=surface|uniform]
"""
import json
import os
import sys
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))

import numpy as np
import torch


def cloud(n, kind, seed=3):
    g = torch.Generator().manual_seed(seed)
    if kind == "uniform":
        return torch.rand(n, 3, generator=g) * 2 - 1
    # a dense object (three ellipsoid surfaces) inside a sparse room shell, like oracle/make_golden_knn.py
    u = torch.nn.functional.normalize(torch.randn(n, 3, generator=g), dim=-1)
    radii = torch.tensor([[0.10, 0.06, 0.05], [0.05, 0.05, 0.08], [0.03, 0.07, 0.03]])
    centre = torch.tensor([[0.0, 0.0, 0.0], [0.08, 0.02, 0.03], [-0.05, -0.04, 0.06]])
    which = torch.randint(0, 3, (n,), generator=g)
    pts = centre[which] + u * radii[which] + 0.002 * torch.randn(n, 3, generator=g)
    shell = torch.rand(n, generator=g) < 0.15
    pts[shell] = 2.0 * torch.nn.functional.normalize(torch.randn(int(shell.sum()), 3, generator=g), dim=-1)
    return pts.float()


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
    k = int(sys.argv[2]) if len(sys.argv) > 2 else 16
    kind = sys.argv[3] if len(sys.argv) > 3 else "surface"
    from fusionsense_b200.knn import KnnIndex, gaussian_density, knn_sk

    x_cpu = cloud(n, kind)
    x = x_cpu.cuda()
    knn_sk(x, x, k)  # warm-up (library load, allocator)
    torch.cuda.synchronize()
    g = torch.Generator().manual_seed(1)
    ls = torch.log(0.004 * torch.exp(0.5 * torch.randn(n, 3, generator=g))).cuda()
    q = torch.randn(n, 4, generator=g).cuda()
    op = torch.randn(n, 1, generator=g).cuda()
    torch.cuda.synchronize()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
    ev[0].record()
    index = KnnIndex(x)
    ev[1].record()
    idx, dist = index.query(index.x, k + 1, drop_first=1, return_distances=True)
    ev[2].record()
    dens = gaussian_density(index.x, idx, index.x, ls, q, op)
    ev[3].record()
    torch.cuda.synchronize()
    ms_build, ms_query, ms_dens = ev[0].elapsed_time(ev[1]), ev[1].elapsed_time(ev[2]), ev[2].elapsed_time(ev[3])
    t0 = time.perf_counter()
    out = knn_sk(x, x, k)
    torch.cuda.synchronize()
    wall = time.perf_counter() - t0
    unresolved = int(index.last_unresolved)

    # the reference's path on the host cores: fit on the whole cloud, query a bounded sample, scale to all queries
    from sklearn.neighbors import NearestNeighbors

    sample = min(n, 50_000)
    t0 = time.perf_counter()
    nn_model = NearestNeighbors(n_neighbors=k + 1, algorithm="auto", metric="euclidean").fit(x_cpu.numpy())
    t_fit = time.perf_counter() - t0
    t0 = time.perf_counter()
    d_ref, i_ref = nn_model.kneighbors(x_cpu[:sample].numpy())
    t_q = time.perf_counter() - t0
    cpu_s = t_fit + t_q * n / sample
    same = float((out[:sample].cpu().numpy() == i_ref[:, 1:]).all(axis=1).mean())
    dmax = float(np.abs(dist[:sample].cpu().numpy() - d_ref[:, 1:]).max())
    print(json.dumps({
        "metric": "knn_queries_per_s", "workload": f"{k}-NN of {n} points against themselves ({kind} cloud), exact",
        "value": n / ((ms_build + ms_query) * 1e-3), "unit": "query/s", "ms_build": ms_build, "ms_query": ms_query,
        "ms_total_wall": wall * 1e3, "grid": index.g, "queries_finished_by_brute_force": unresolved,
        "density": {"ms": ms_dens, "samples": n, "k": k, "gbs_algorithmic": n * (12 + k * (8 + 44) + 4) / (ms_dens * 1e-3) / 1e9},
        "cpu_baseline": {"kind": "reference", "what": "dn_splatter/utils/knn.py knn_sk = sklearn NearestNeighbors "
                         "(kd-tree) on the host", "cores": os.cpu_count(), "fit_s": t_fit,
                         "sample": f"{sample} of {n} queries, scaled", "value": n / cpu_s, "unit": "query/s", "total_s": cpu_s},
        "speedup_vs_cpu": cpu_s / ((ms_build + ms_query) * 1e-3),
        "rows_identical_to_sklearn_on_sample": same, "max_abs_distance_difference": dmax,
    }))


if __name__ == "__main__":
    main()
