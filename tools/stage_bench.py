"""Per-stage kernel times of one DN-Splatter step on the other BASELINE.json configurations (not the bench line):

  python tools/stage_bench.py cfg1|cfg2|cfg4 [steps]

CUDA events around each C-ABI call on the launching stream (ops.kernel_timer), warm caches, after warm-up; prints
one JSON object with the mean milliseconds per stage, n_isects, and raster fwd / fwd+bwd Mpix/s.
"""
import json
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch

from fusionsense_b200 import ops
from fusionsense_b200.dn_step import DNSplatterStep, DNSplatterStepConfig
from fusionsense_b200.gsplat.cuda_legacy import _wrapper as legacy
from fusionsense_b200.synthetic import make_scene

CFGS = {
    "cfg1": dict(n=50_000, w=640, h=480, kind="random", views=3, cfg_id=1),
    "cfg2": dict(n=300_000, w=640, h=480, kind="bunny", views=9, cfg_id=2),
    "cfg4": dict(n=1_000_000, w=1920, h=1080, kind="random", views=8, cfg_id=4),
    "cfg5": dict(n=3_000_000, w=3840, h=2160, kind="random", views=4, cfg_id=5),
}


def main():
    name = sys.argv[1] if len(sys.argv) > 1 else "cfg4"
    steps = int(sys.argv[2]) if len(sys.argv) > 2 else 10
    c = CFGS[name]
    scene = make_scene(c["n"], c["w"], c["h"], n_views=c["views"], cfg_id=c["cfg_id"], kind=c["kind"])
    model = DNSplatterStep(scene, DNSplatterStepConfig(), device="cuda", step=3000)
    targets = {v: model.render_targets(v) for v in range(c["views"])}
    for i in range(4):
        model.train_iteration(i % c["views"], targets[i % c["views"]])
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for i in range(steps):
        model.train_iteration(i % c["views"], targets[i % c["views"]])
    e.record()
    torch.cuda.synchronize()
    ms_step = s.elapsed_time(e) / steps
    # per-stage times in a second leg: a spin kernel ahead of every start event keeps launch latency out of them
    with ops.kernel_timer.collect(pad_cycles=400_000):
        for i in range(steps):
            model.train_iteration(i % c["views"], targets[i % c["views"]])
        kt = ops.kernel_timer.summary()
    P = c["w"] * c["h"]
    # fused_passes (default): one kernel pair composites RGB + depth and the normals ("D4+3")
    fwd = kt.get("raster_fwd_D4+3", kt.get("raster_fwd_D4", (float("nan"),)))[0]
    bwd = kt.get("raster_bwd_D4+3", kt.get("raster_bwd_D4", (float("nan"),)))[0]
    out = {
        "config": name, **c, "steps": steps, "ms_per_step": ms_step, "iter_per_s": 1e3 / ms_step,
        "n_isects": int(legacy._LAST_BINNING.get("n_isects", 0)), "n_visible": int((model.radii > 0).sum()),
        "kernel_ms": {k: round(v[0], 4) for k, v in sorted(kt.items())},
        "raster_fwd_mpix_per_s": P / (fwd * 1e-3) / 1e6, "raster_fwd_bwd_mpix_per_s": P / ((fwd + bwd) * 1e-3) / 1e6,
    }
    print(json.dumps(out))


if __name__ == "__main__":
    main()
