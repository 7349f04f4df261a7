"""Repeat the 1M-point 16-NN self query a few times in one process (CUDA events per repetition): separates first-use
and clock-ramp effects from the steady-state time.  python tools/knn_repeat.py [uniform|surface]"""
import json
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tools"))
import torch

from knn_bench import cloud


def main():
    kind = sys.argv[1] if len(sys.argv) > 1 else "uniform"
    from fusionsense_b200.knn import KnnIndex

    x = cloud(1_000_000, kind).cuda()
    out = {"kind": kind, "build_ms": [], "query_ms": [], "query_dist_ms": []}
    for rep in range(6):
        e = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
        e[0].record()
        index = KnnIndex(x)
        e[1].record()
        index.query(index.x, 17, drop_first=1)
        e[2].record()
        index.query(index.x, 17, drop_first=1, return_distances=True)
        e[3].record()
        torch.cuda.synchronize()
        out["build_ms"].append(round(e[0].elapsed_time(e[1]), 3))
        out["query_ms"].append(round(e[1].elapsed_time(e[2]), 3))
        out["query_dist_ms"].append(round(e[2].elapsed_time(e[3]), 3))
    out["unresolved"] = int(index.last_unresolved)
    print(json.dumps(out))


if __name__ == "__main__":
    main()
