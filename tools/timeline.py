"""GPU timeline of the bench step from torch.profiler (kineto): busy time, idle gaps and what the host was doing.

  python tools/timeline.py [out.json]     (run on the GPU box; writes a compact JSON of kernel intervals)
"""
import json
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch
from torch.profiler import ProfilerActivity, profile

import bench

out = sys.argv[1] if len(sys.argv) > 1 else "gpurun_out/timeline.json"
model = bench.build_model("cfg2", torch.device("cuda", 0))
targets = {v: model.render_targets(v) for v in range(bench.CONFIGS["cfg2"]["views"])}
for i in range(8):
    model.train_iteration(i % 9, targets[i % 9])
torch.cuda.synchronize()
STEPS = 6
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
    for i in range(STEPS):
        model.train_iteration(i % 9, targets[i % 9])
    torch.cuda.synchronize()
trace = "/tmp/trace.json"
prof.export_chrome_trace(trace)
ev = json.load(open(trace))["traceEvents"]
kern = sorted([(e["ts"], e["dur"], e["name"][:60]) for e in ev if e.get("cat") in ("kernel", "gpu_memcpy", "gpu_memset")])
cpu = sorted([(e["ts"], e["dur"], e["name"][:60]) for e in ev if e.get("cat") in ("cpu_op", "cuda_runtime", "python_function", "user_annotation") and e.get("dur", 0) > 0])
json.dump({"steps": STEPS, "kernels": kern, "cpu": cpu}, open(out, "w"))
t0, t1 = kern[0][0], kern[-1][0] + kern[-1][1]
busy = sum(d for _, d, _ in kern)
print(f"steps {STEPS}: wall {t1 - t0:.0f} us, gpu busy {busy:.0f} us ({100 * busy / (t1 - t0):.1f}%), per step wall {(t1 - t0) / STEPS:.0f} busy {busy / STEPS:.0f}")
gaps = []
for (a, da, na), (b, db, nb) in zip(kern[:-1], kern[1:]):
    g = b - (a + da)
    if g > 8:
        gaps.append((g, na, nb))
gaps.sort(reverse=True)
print("largest idle gaps (us, after kernel -> before kernel):")
for g, na, nb in gaps[:25]:
    print(f"  {g:8.1f}  {na}  ->  {nb}")
print("total gap time in gaps > 8us:", sum(g for g, _, _ in gaps), "count", len(gaps))
small = sum(max(0, b - (a + da)) for (a, da, _), (b, _, _) in zip(kern[:-1], kern[1:]) if b - (a + da) <= 8)
print("total gap time in gaps <= 8us:", small)
