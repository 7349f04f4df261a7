"""Multi-GPU check of the training path, run under torchrun on a B200 box:

  python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 tools/train_multi_gpu_check.py [peer|nccl]

Every rank holds a full replica of a 40k-Gaussian scene, renders its own view per iteration through the captured step
(views sharded r::world, dist.shard_views), exchanges gradients (our NVLink kernels or the captured NCCL all-reduce),
and at a refine step runs dist.synchronised_refinement (SUM / SUM / MAX of the densification statistics on NCCL, shared
split RNG).  Asserted: replicas bit-identical before the refine step, after it, and after further captured steps on
the new Gaussian count; the Gaussian count changed.  One JSON line on rank 0."""
import json
import os
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch
import torch.distributed as dist


def main():
    mode = sys.argv[1] if len(sys.argv) > 1 else "peer"
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", rank)))
    dev = torch.device("cuda", torch.cuda.current_device())
    dist.init_process_group("nccl", device_id=dev)
    from fusionsense_b200 import dist as fdist
    from fusionsense_b200.dn_step import DNSplatterStep, DNSplatterStepConfig
    from fusionsense_b200.graph_step import GraphedDNSplatterStep
    from fusionsense_b200.synthetic import make_scene

    n_views = 8
    scene = make_scene(40_000, 320, 240, n_views=n_views, cfg_id=97, kind="bunny", fx=300.0)
    model = DNSplatterStep(scene, DNSplatterStepConfig(), device=dev, step=3190)
    targets = {v: model.render_targets(v) for v in range(n_views)}
    sync = fdist.PeerGradExchange() if mode == "peer" else fdist.GradSync()
    runner = GraphedDNSplatterStep(model, targets, grad_sync=sync, loss_scale=1.0 / world)
    params = list(model.gauss_params.values())
    out = {"mode": mode, "world": world, "n0": model.num_points}

    def steps(k):
        for _ in range(k):
            v = fdist.shard_views(model.step, rank, world, n_views)[0]
            runner.train_iteration(v)
        info = runner.poll()
        assert info["overflowed_steps"] == 0, info
        return info["loss"]

    out["loss_a"] = steps(10)  # 3190 .. 3199
    out["identical_before_refine"] = fdist.replicas_identical(params)
    assert model.step == 3200
    fdist.synchronised_refinement(model, model.optimizers, model.step, seed=1234)
    params = list(model.gauss_params.values())
    out["n1"] = model.num_points
    out["identical_after_refine"] = fdist.replicas_identical(params)
    out["loss_b"] = steps(6)  # re-captured on the new Gaussian count
    params = list(model.gauss_params.values())
    out["identical_after_more_steps"] = fdist.replicas_identical(params)
    out["captures"] = runner.captures
    out["graph_launches_per_step"] = 1 if runner.graph_tail is None else 2
    ok = (out["identical_before_refine"] and out["identical_after_refine"] and out["identical_after_more_steps"]
          and out["n1"] != out["n0"] and out["captures"] >= 2)  # + 1 when the binary-opacity window opens (step 3201)
    out["ok"] = bool(ok)
    if rank == 0:
        print(json.dumps(out), flush=True)
    dist.barrier()
    torch.cuda.synchronize()
    sys.stdout.flush()
    # no destroy_process_group: tearing down a communicator whose collectives sit in (re-)captured CUDA graphs hung the
    # r02g run until its timeout; the process is done, leave at once
    os._exit(0 if ok else 1)


if __name__ == "__main__":
    main()
