"""cfg3 of BASELINE.json: VisualHull voxel carving, 512^3 voxels from 9 masks 640x480.

  python tools/hull_bench.py [n_per_axis=512] [reps=5]          # one GPU
  torchrun --nproc-per-node N tools/hull_bench.py 512 5          # voxel-slab sharded (z slabs, §8e)

Times votes / count+scan / compaction with CUDA events (warm, after one untimed pass), reports each kernel against
the bound DESIGN.md §4 names for it, and times the CPU oracle (oracle/visual_hull_ref.py, the numpy restatement of
utils/VisualHull.py:149-191) on a bounded sample of z-planes of the same grid.  One JSON line on stdout (rank 0).
The masks / cameras are the committed golden capture (tests/golden/visual_hull_201.npz): 9 views, 640x480.
"""
import json
import os
import sys
import tempfile
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))

import numpy as np
import torch


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 512
    reps = int(sys.argv[2]) if len(sys.argv) > 2 else 5
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        import torch.distributed as dist

        dist.init_process_group("nccl", device_id=dev)

    from fusionsense_b200 import visual_hull as vh
    from fusionsense_b200._abi import lib
    from tests.golden_io import load_visual_hull_golden, write_visual_hull_capture

    g = load_visual_hull_golden()
    with tempfile.TemporaryDirectory() as td:
        path = write_visual_hull_capture(td, g)
        mats, centre, names = vh.read_hull_cameras(path)
        masks = vh.read_masks(path, names)
    xs, ys, zs = vh.hull_grid(centre, half_extent=0.5, n_per_axis=n)
    M, H, W = masks.shape
    carver = vh.HullCarver(mats, masks, xs, ys, zs, device=dev, rank=rank, world_size=world)

    def one_pass(timed):
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
        if timed:
            torch.cuda._sleep(400_000)  # the host runs ahead: events and kernels queue back to back
        ev[0].record()
        maxv = carver.vote()  # 8-byte D2H of the slab maximum (a host sync, as in the product path)
        ev[1].record()
        if world > 1:
            maxv = vh.reduce_max(maxv, dev)
        iso = vh.iso_value(maxv, 5)
        ev[2].record()
        pts = carver.extract(iso)  # count + scan (+ 8-byte D2H of n_occ) + compaction
        ev[3].record()
        torch.cuda.synchronize()
        return maxv, iso, pts, (ev[0].elapsed_time(ev[1]), ev[2].elapsed_time(ev[3]), ev[0].elapsed_time(ev[3]))

    maxv, iso, pts, _ = one_pass(False)
    n_occ_local = int(pts.shape[0])
    ts = []
    l0 = lib.fsb_launch_count()
    for _ in range(reps):
        if world > 1:
            dist.barrier()
        ts.append(one_pass(True)[3])
    launches = lib.fsb_launch_count() - l0
    t = torch.tensor(ts, device=dev).min(dim=0).values  # best of reps per stage
    n_occ = torch.tensor([n_occ_local], device=dev, dtype=torch.int64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dist.all_reduce(n_occ)
    votes_ms, extract_ms, total_ms = (float(v) for v in t)
    V = n ** 3
    Vr = carver.V
    peaks = json.loads((ROOT / "MEASURED_PEAKS.json").read_text()) if (ROOT / "MEASURED_PEAKS.json").exists() else {}
    hbm = float(peaks.get("hbm_gbs", 6650.0))
    # votes: 8 B written per voxel (+ masks once); per (voxel, view): 11 DFMA/DMUL + 2 IEEE fp64 divisions + 2 floors
    # + 2 DADD = 17 fp64 results (a division counted once).  FP64 peak: 148 SM x 64 lanes x 2 x 1.965 GHz (nominal).
    votes_bytes = Vr * 8 + M * H * W
    fp64_flop = Vr * M * (2 * 11 + 2 + 2 + 2)
    fp64_peak = 148 * 64 * 2 * 1.965e9 / 1e12
    # count re-reads the votes, compaction re-reads them and writes 24 B (+8 B index) per occupied voxel
    extract_bytes = Vr * 8 * 2 + n_occ_local * 24
    line = {
        "metric": "visual_hull_carve_voxels_per_s", "value": V / (total_ms * 1e-3), "unit": "voxel/s",
        "n_gpus": world, "ms_total": total_ms, "scaling": "strong",
        "config": {"workload": f"cfg3: VisualHull {n}^3 voxels, {M} masks {W}x{H}, error=5, z-slab sharded x{world}",
                   "voxels": V, "voxels_per_rank": Vr, "n_occupied": int(n_occ), "maxv": maxv, "iso": iso},
        "stages_ms": {"votes(+8B D2H max)": votes_ms, "count+scan+compact": extract_ms},
        "roofline": {
            "votes": {"bound": "fp64 pipe (IEEE divisions) / HBM write", "algorithmic_bytes": votes_bytes,
                      "gbs": votes_bytes / (votes_ms * 1e-3) / 1e9, "hbm_frac": votes_bytes / (votes_ms * 1e-3) / 1e9 / hbm,
                      "fp64_flop": fp64_flop, "fp64_tflops": fp64_flop / (votes_ms * 1e-3) / 1e12,
                      "fp64_peak_tflops_nominal": fp64_peak,
                      "fp64_frac": fp64_flop / (votes_ms * 1e-3) / 1e12 / fp64_peak},
            "extract": {"bound": "hbm", "algorithmic_bytes": extract_bytes,
                        "gbs": extract_bytes / (extract_ms * 1e-3) / 1e9,
                        "hbm_frac": extract_bytes / (extract_ms * 1e-3) / 1e9 / hbm},
            "hbm_peak_gbs": hbm,
        },
        "gpu_launches": int(launches),
    }
    if rank == 0 and world == 1 and os.environ.get("FSB_NO_CPU") != "1":
        from oracle import visual_hull_ref as ref

        planes = 4
        t0 = time.perf_counter()
        ref.project_votes(mats, masks, xs, ys, zs[:planes])
        dt = time.perf_counter() - t0
        per_voxel = dt / (planes * n * n)
        line["cpu_baseline"] = {"value": 1.0 / per_voxel, "unit": "voxel/s", "cores": os.cpu_count(), "kind": "port",
                                "sample": f"{planes} z-planes ({planes * n * n} voxels) of the same {n}^3 grid through "
                                          f"oracle/visual_hull_ref.project_votes (numpy float64 matmul + gather, "
                                          f"votes only), {dt:.2f} s; full grid extrapolates to {per_voxel * V:.0f} s"}
    if rank == 0:
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
