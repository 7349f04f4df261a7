"""cProfile of the host side of the bench step (where do the non-GPU milliseconds go)."""
import cProfile, pstats, sys, io
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch
import bench

model = bench.build_model("cfg2", torch.device("cuda", 0))
targets = {v: model.render_targets(v) for v in range(bench.CONFIGS["cfg2"]["views"])}
for i in range(5):
    model.train_iteration(i % 9, targets[i % 9])
torch.cuda.synchronize()
pr = cProfile.Profile()
pr.enable()
for i in range(30):
    model.train_iteration(i % 9, targets[i % 9])
torch.cuda.synchronize()
pr.disable()
for key in ("tottime", "cumtime"):
    s = io.StringIO()
    pstats.Stats(pr, stream=s).sort_stats(key).print_stats(45)
    print(s.getvalue())
