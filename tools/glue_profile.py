"""Which torch (non-libfsb200) kernels run inside one eager DN-Splatter step, and which Python line launches them:

  python tools/glue_profile.py cfg4 > gpurun_out/glue_cfg4.txt

torch.profiler with stacks over three iterations; prints, per at:: kernel, launches / iteration, microseconds /
iteration and the innermost frames of this repository on the stack.  libfsb200 kernels are listed once at the end."""
import collections
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch
from torch.profiler import ProfilerActivity, profile

from fusionsense_b200.dn_step import DNSplatterStep, DNSplatterStepConfig
from fusionsense_b200.synthetic import make_scene
from tools.stage_bench import CFGS


def main():
    name = sys.argv[1] if len(sys.argv) > 1 else "cfg4"
    c = CFGS[name]
    scene = make_scene(c["n"], c["w"], c["h"], n_views=c["views"], cfg_id=c["cfg_id"], kind=c["kind"])
    model = DNSplatterStep(scene, DNSplatterStepConfig(), device="cuda", step=3000)
    targets = {v: model.render_targets(v) for v in range(c["views"])}
    for i in range(4):
        model.train_iteration(i % c["views"], targets[i % c["views"]])
    torch.cuda.synchronize()
    iters = 3
    with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA], with_stack=True) as prof:
        for i in range(iters):
            model.train_iteration(i % c["views"], targets[i % c["views"]])
        torch.cuda.synchronize()
    agg = collections.OrderedDict()
    for ev in prof.events():
        if ev.device_type != torch.autograd.DeviceType.CUDA:
            continue
        a = agg.setdefault(ev.name[:70], [0, 0.0])
        a[0] += 1
        a[1] += ev.device_time if hasattr(ev, "device_time") else ev.cuda_time
    print(f"# {name}: device kernels over {iters} eager iterations")
    for k, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"{a[1] / iters:9.1f} us/it {a[0] / iters:5.1f} /it  {k}")
    print("\n# CPU ops with stacks (top by device time), repository frames only")
    rows = prof.key_averages(group_by_stack_n=12)
    rows = sorted(rows, key=lambda r: -(getattr(r, "device_time_total", 0) or getattr(r, "cuda_time_total", 0)))
    shown = 0
    for r in rows:
        dt = getattr(r, "device_time_total", 0) or getattr(r, "cuda_time_total", 0)
        if dt <= 0 or not r.key.startswith("aten::"):
            continue
        frames = [f for f in (r.stack or []) if "/fusionsense_b200/" in f or "/tests/" in f]
        print(f"{dt / iters:9.1f} us/it {r.count / iters:5.1f} /it  {r.key:32s} {' <- '.join(f.split('/repo/')[-1] for f in frames[:3])}")
        shown += 1
        if shown >= 40:
            break


if __name__ == "__main__":
    main()
