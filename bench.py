#!/usr/bin/env python
"""bench.py — DN-Splatter train step throughput on B200 (BASELINE.json metric, configs[1]).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

Workload (config.workload = "cfg2"): FusionSense sparse-view DN-Splatter training, 9 views 640x480
RealSense-shaped synthetic bunny scene, 300k Gaussians, one camera view per GPU per iteration:
get_outputs (RGB+ED rasterization + legacy normals pass) -> losses -> backward -> Adam -> after_train.
One JSON line on stdout (rank 0).  See DESIGN.md §Measurement for the definition of every key.

--impl reference: the same step through the CPU oracle (oracle/gsplat_ref.py standing in for gsplat, which is
not vendored in the reference tree) on the box's host cores, each step a bounded sample of the workload.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

N_GAUSS = 300_000
WIDTH, HEIGHT = 640, 480
N_VIEWS = 9
WORKLOAD = "cfg2: DN-Splatter train step, 9 views 640x480 synthetic bunny, 300k Gaussians, 1 view/GPU/iter"
METRIC = "dn_splatter_train_iter_per_s"
UNIT = "iter/s"


def _peaks():
    p = ROOT / "MEASURED_PEAKS.json"
    if p.exists():
        d = json.loads(p.read_text())
        return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def _ncu_traffic():
    """DRAM bytes (dram__bytes_read.sum + dram__bytes_write.sum) of ONE raster_bwd_kernel<4> launch on this workload,
    from the committed `ncu --set full` capture (profiles/raster_bwd_traffic.json, written by
    tools/ncu_traffic.py from the raw page); None when no capture is committed."""
    p = ROOT / "profiles" / "raster_bwd_traffic.json"
    if not p.exists():
        return None, None
    d = json.loads(p.read_text())
    return d.get("dram_bytes_per_launch"), d.get("source")


def _fp32_roofline(pairs, ktimes, D, clocks):
    """Compositing is bound by the FP32 pipe, so next to the HBM fraction BASELINE.json asks for: algorithmic
    pair-flops (forward Q (14 + 2 D), backward Q (40 + 6 D), Q = blended (pixel, entry) pairs counted
    on the device by fsb_raster_pair_count) over the kernels' CUDA-event times, against 148 SMs x 128 FP32 lanes x
    2 flop x the SM clock sampled under load."""
    q = pairs.get(f"D{D}")
    if not q:
        return None
    mhz = (clocks or {}).get("sm_max_mhz") or 1965.0
    peak = 148 * 128 * 2 * mhz * 1e6 / 1e12
    out = {"pairs_blended": q["blended"], "pairs_visited": q["visited"], "peak_tflops": peak,
           "peak_source": f"148 SM x 128 lanes x 2 x {mhz:.0f} MHz (nominal FP32 FMA rate, not measured)"}
    for name, flop_per_pair in (("raster_fwd", 14 + 2 * D), ("raster_bwd", 40 + 6 * D)):
        ms = ktimes.get(f"{name}_D{D}", (float("nan"), 0))[0]
        if ms == ms and ms > 0:
            tf = q["blended"] * flop_per_pair / (ms * 1e-3) / 1e12
            out[name] = {"flop_per_pair": flop_per_pair, "achieved_tflops": tf, "frac": tf / peak, "kernel_ms": ms}
    return out


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 500 ms from before the warm-up on; every sample is stamped on
    arrival and `stop(t0, t1)` reports the ones that fall inside the timed region (all of them, flagged, if the region
    was too short to catch one).  The period is deliberately long: every nvidia-smi query stalls the GPU for ~3.6 ms
    (r01l: five samples inside a 219 ms region cost 8 % of the measured rate), so the default run times 1000 steps
    (~1 s) and takes two samples inside it."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.rows, self.proc, self.gpu = [], None, gpu_index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "500", "-i", str(self.gpu)], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.perf_counter(), [c.strip() for c in line.split(",")]))

    def stop(self, t0=None, t1=None):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.1)
        self.proc.terminate()
        rows = list(self.rows)
        have = t0 is not None and t1 is not None
        inside = [r for t, r in rows if have and t0 <= t <= t1 + 0.05]
        # a region shorter than the sampling period: the samples next to it (warm-up steps before, e2e leg after —
        # the same workload) stand in
        near = [r for t, r in rows if have and t0 - 0.3 <= t <= t1 + 0.3]
        window = "timed region" if inside else ("timed region +- 0.3 s (warm-up / e2e leg)" if near else "whole run")
        sm, mx, reasons = [], [], set()
        for r in (inside or near or [r for _, r in rows]):
            try:
                sm.append(float(r[1])); mx.append(float(r[2]))
            except (ValueError, IndexError):
                continue
            for name, idx in (("hw_slowdown", 5), ("hw_thermal_slowdown", 6), ("sw_thermal_slowdown", 7),
                              ("sw_power_cap", 8)):
                if len(r) > idx and r[idx].lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm),
                "window": window}


def build_model(device, gsplat_module=None, fused=True):
    import torch
    from fusionsense_b200.dn_step import DNSplatterStep, DNSplatterStepConfig
    from fusionsense_b200.synthetic import make_scene

    scene = make_scene(N_GAUSS, WIDTH, HEIGHT, n_views=N_VIEWS, cfg_id=2, kind="bunny")
    cfg = DNSplatterStepConfig(fused_optimizer=fused)
    if not fused:
        cfg.stop_split_at = 0  # the CPU oracle has no absgrad side channel; after_train is skipped there
    return DNSplatterStep(scene, cfg, device=device, step=3000, gsplat_module=gsplat_module)


# ---------------------------------------------------------------------------------------------
# reference arm: CPU oracle
# ---------------------------------------------------------------------------------------------
def cpu_step_seconds(steps: int, warmup: int, sample_tiles_frac: float = 1.0):
    """Time the oracle-driven train step on the host cores. Returns (seconds per full step, cores, sample text)."""
    import torch
    from oracle import gsplat_ref as ref

    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    model = build_model("cpu", gsplat_module=ref, fused=False)
    targets = {0: model.render_targets(0)}
    ts = []
    for i in range(warmup + steps):
        t0 = time.perf_counter()
        model.train_iteration(0, targets[0])
        dt = time.perf_counter() - t0
        if i >= warmup:
            ts.append(dt)
    sample = (f"{steps} full train step(s) (RGB+ED pass, normals pass, losses, backward, torch Adam) of the same "
              f"300k-Gaussian 640x480 scene through oracle/gsplat_ref.py, torch CPU fp32, {cores} threads")
    return statistics.median(ts), cores, sample


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    steps = max(1, min(args.steps, 3))
    sec, cores, sample = cpu_step_seconds(steps, warmup=min(args.warmup, 1))
    v = 1.0 / sec
    line = {
        "impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": steps,
        "warmup": min(args.warmup, 1), "ms_per_step": sec * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": WORKLOAD, "gaussians": N_GAUSS, "width": WIDTH, "height": HEIGHT,
                   "note": "gsplat==1.0.0 is not vendored in the reference tree and not installed; the reference's "
                           "CPU path is its algorithm restated in oracle/ (kind=port)"},
        "cpu_baseline": {"value": v, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    _emit(line)


# ---------------------------------------------------------------------------------------------
# our arm
# ---------------------------------------------------------------------------------------------
def run_ours(args):
    import torch
    import torch.distributed as dist

    from fusionsense_b200 import ops
    from fusionsense_b200._abi import lib

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py (impl=ours) needs a CUDA device: fusionsense_b200 has no CPU fallback")
    torch.cuda.set_device(local_rank)
    device = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=device)

    model = build_model(device)
    views = list(range(N_VIEWS))
    dev_targets = {v: model.render_targets(v) for v in views}
    host_targets = {v: {k: t.cpu().pin_memory() for k, t in d.items()} for v, d in dev_targets.items()}
    h2d_bytes = sum(t.numel() * t.element_size() for t in host_targets[0].values())

    params = [model.gauss_params[k] for k in model.config.lrs]
    graph_mode = args.mode == "graph"

    from fusionsense_b200.dist import GradSync

    # one flat NCCL all-reduce (SUM) over all Gaussian parameter gradients (59 floats per Gaussian) plus the
    # overflow flag of the static-capacity step; the loss is pre-scaled by 1/world, so the sum is the mean
    allreduce_grads = GradSync()

    runner = None
    if graph_mode:
        from fusionsense_b200.graph_step import GraphedDNSplatterStep

        runner = GraphedDNSplatterStep(model, dev_targets, grad_sync=allreduce_grads if world > 1 else None,
                                       loss_scale=1.0 / world)

    def eager_step(i, batch):
        v = (i * world + rank) % N_VIEWS
        for opt in model.optimizers.values():
            opt.zero_grad(set_to_none=True)
        outputs = model.get_outputs(v)
        loss = model.get_loss_dict(outputs, batch(v))["main_loss"]
        (loss / world if world > 1 else loss).backward()
        if world > 1:
            allreduce_grads(params)
        model.optimizers["means"].param_groups[0]["lr"] = model._means_lr()
        model.optimizer_step()
        model.after_train()
        model.step += 1
        return loss

    def resident(v):
        return dev_targets[v]

    def from_host(v):
        return {k: t.to(device, non_blocking=True) for k, t in host_targets[v].items()}

    def one_step(i, staged, read_loss, n_total=1 << 30):
        """One training iteration; `staged`: this step's targets come from pinned host memory; `read_loss`: the
        step's result is read back to the host."""
        if runner is None:
            loss = eager_step(i, from_host if staged else resident)
            if read_loss:
                float(loss)
            return
        v = (i * world + rank) % N_VIEWS
        if staged and i == 0:
            runner.stage_async(v, host_targets[v])
        runner.train_iteration(v)  # waits for this view's copy
        if staged and i + 1 < n_total:
            # the next step's inputs travel while this step's kernels run (copy stream, pinned source)
            vn = ((i + 1) * world + rank) % N_VIEWS
            runner.stage_async(vn, host_targets[vn])
        if read_loss:
            # 32-byte D2H read of [loss, overflow count, n_isects x2] per step; the host consumes step i - 1's
            # values here and the last step's after the loop (timed() drains the ring before the closing event)
            runner.read_result_async()

    def timed(n_steps, staged, read_loss):
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        for i in range(n_steps):
            one_step(i, staged, read_loss, n_steps)
        if read_loss and runner is not None:
            runner.poll()  # the last step's result (synchronises)
        e.record()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        ms = torch.tensor([s.elapsed_time(e)], device=device)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms.item())

    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    # warm-up (also primes the caching allocator and, in graph mode, captures the step)
    for i in range(args.warmup):
        one_step(i, False, False)
    torch.cuda.synchronize()
    if runner is not None and runner.poll()["new_overflows"]:
        for i in range(args.warmup):  # capacity grew: capture again before timing
            one_step(i, False, False)
        torch.cuda.synchronize()

    launches0 = lib.fsb_launch_count()
    replays0 = runner.replays if runner else 0
    profiling = os.environ.get("FSB_PROFILE") == "1"  # ncu --profile-from-start off: capture the timed steps only
    if profiling:
        torch.cuda.profiler.start()
    t_region0 = time.perf_counter()
    ms_total = timed(args.steps, False, False)
    t_region1 = time.perf_counter()
    if profiling:
        torch.cuda.profiler.stop()
    launches = lib.fsb_launch_count() - launches0
    graph_info = None
    if runner is not None:
        info = runner.poll()
        if info["new_overflows"]:
            raise SystemExit(f"bench: {info['new_overflows']} timed step(s) overflowed the intersection capacity; rerun")
        # replays launch the captured kernels without passing through the library's entry points
        launches += (runner.replays - replays0) * runner.launches_per_replay
        graph_info = {"captures": runner.captures, "capacity": runner.capacity, "n_isects": info["n_isects"],
                      "n_isects_normals": info["n_isects_normals"], "libfsb200_launches_per_replay": runner.launches_per_replay}
    clocks = sampler.stop(t_region0, t_region1) if rank == 0 else None
    ms_e2e = timed(args.steps, True, True)
    if runner is not None and runner.poll()["overflowed_steps"]:
        raise SystemExit("bench: a step of the e2e leg overflowed the intersection capacity; rerun")

    # per-kernel CUDA-event times of the same workload: eager launches (a replayed graph offers no place to record
    # events between its kernels), same kernels, same sizes, same stream, after the timed legs
    with ops.kernel_timer.collect(pad_cycles=400_000):
        for i in range(6):
            eager_step(i, resident)
        ktimes = ops.kernel_timer.summary()
    # pair counts (Q of SURVEY.md §8d) of one more eager step, counted by fsb_raster_pair_count after each forward
    ops.pair_probe.enabled = True
    eager_step(5, resident)
    torch.cuda.synchronize()
    ops.pair_probe.enabled = False
    pairs = ops.pair_probe.summary()

    ms_per_step = ms_total / args.steps
    value = world * 1e3 / ms_per_step
    e2e_value = world * args.steps * 1e3 / ms_e2e

    # roofline of the dominant kernel: raster backward of the RGB+ED pass (D = 4)
    from fusionsense_b200.gsplat.cuda_legacy import _wrapper as _legacy

    I = int(_legacy._LAST_BINNING.get("n_isects", 0))  # intersections of the last timed step
    Nv = int((model.radii > 0).sum().item())
    P = WIDTH * HEIGHT
    D = 4
    bwd_ms = ktimes.get("raster_bwd_D4", (float("nan"), 0))[0]
    bwd_bytes = I * (28 + 4 * D) + P * (4 * D + 12) + Nv * (32 + 4 * D)
    peak, peak_src = _peaks()
    achieved = bwd_bytes / (bwd_ms * 1e-3) / 1e9 if bwd_ms == bwd_ms and bwd_ms > 0 else None
    traffic, traffic_src = _ncu_traffic()
    # every timed stage against the bound DESIGN.md §4 names for it: algorithmic bytes (SURVEY.md §8d formulas with
    # this step's N, Nv, I, P, K) over the stage's mean CUDA-event time
    N, K, C = N_GAUSS, 16, 1
    tiles = ((WIDTH + 15) // 16) * ((HEIGHT + 15) // 16)
    key_bytes = -(-(32 + max(1, (tiles - 1).bit_length())) // 8)
    stage_bytes = {
        "project_sh_fwd": C * N * 68 + Nv * (12 * K + 12),
        "project_sh_bwd": C * N * 4 + Nv * (76 + 12 * K) + N * (40 + 12 * K),
        "radix_sort": I * 8 + I * 24 * key_bytes,
        "adam_multi": sum(p.numel() for p in params) * 28,
        "raster_fwd_D4": I * (28 + 16) + P * (16 + 8), "raster_fwd_D3": I * (28 + 12) + P * (12 + 8),
        "raster_bwd_D4": bwd_bytes, "raster_bwd_D3": I * (28 + 12) + P * (12 + 12) + Nv * (32 + 12),
    }
    stages = {}
    for name, nbytes in stage_bytes.items():
        ms = ktimes.get(name, (float("nan"), 0))[0]
        if ms == ms and ms > 0:
            gbs = nbytes / (ms * 1e-3) / 1e9
            stages[name] = {"algorithmic_bytes": int(nbytes), "ms": round(ms, 4), "gbs": round(gbs, 1),
                            "hbm_frac": round(gbs / peak, 4),
                            "bound": "fp32" if name.startswith("raster") else "hbm"}
    roofline = {
        "kernel": "raster_bwd_kernel<4> (RGB+ED pass)", "bound": "hbm", "achieved": achieved, "peak": peak,
        "unit": "GB/s", "frac": (achieved / peak) if achieved else None, "traffic": traffic,
        "traffic_source": traffic_src,
        "peak_source": peak_src, "algorithmic_bytes": bwd_bytes, "kernel_ms": bwd_ms,
        "n_isects": I, "n_visible": Nv,
        "note": "compositing is FP32/MUFU-bound, not HBM-bound (SURVEY.md §8d); the HBM fraction is reported as "
                "BASELINE.json asks, the pipe utilisation is in profiles/",
        "fp32": _fp32_roofline(pairs, ktimes, D, clocks if rank == 0 else None),
        "kernel_ms_all": {k: round(v[0], 4) for k, v in sorted(ktimes.items())},
        "stages": stages,
        "raster_fwd_bwd_mpix_per_s": (P / ((ktimes.get("raster_fwd_D4", (0, 0))[0] + bwd_ms) * 1e-3) / 1e6)
        if bwd_ms == bwd_ms and bwd_ms > 0 else None,
    }

    line = None
    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": WORKLOAD, "gaussians": N_GAUSS, "width": WIDTH, "height": HEIGHT,
                       "views": N_VIEWS, "views_per_iter_per_gpu": 1, "global_views_per_iter": world,
                       "execution": ("one CUDA graph replay per iteration (static-capacity intersection lists, "
                                     "no host sync); every kernel of the eager step runs in every replay"
                                     if graph_mode else "eager launches"),
                       "graph": graph_info,
                       "fused_outputs": bool(model.config.fused_outputs),  # dn_step.py: activations / SH concat /
                       # image glue inside our kernels (True) or as the reference's torch ops (False)
                       "l2": "per-step working set (parameters, Adam state, gradients, intersection lists, images: "
                             ">400 MB touched per step) exceeds the 126 MB L2; no explicit flush"},
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d_bytes,
                    "d2h_bytes_per_step": 32 if graph_mode else 4},
            "gpu_launches": int(launches),
            "gpu_launches_per_step": launches / args.steps,
            "clocks": clocks,
            "roofline": roofline,
        }
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    return line


_REAL_STDOUT = None


def _quiet_stdout():
    """stdout must carry exactly one JSON line: NCCL prints its version banner to fd 1 at communicator creation
    (seen in the N=2 run), so everything written to fd 1 during the run is sent to stderr and the line itself goes
    to the saved descriptor."""
    global _REAL_STDOUT
    sys.stdout.flush()
    _REAL_STDOUT = os.dup(1)
    os.dup2(2, 1)


def _emit(line: dict):
    data = (json.dumps(line) + "\n").encode()
    sys.stdout.flush()
    if _REAL_STDOUT is None:
        os.write(1, data)
    else:
        os.write(_REAL_STDOUT, data)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=1000)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--mode", default="graph", choices=["graph", "eager"],
                    help="graph: the iteration is one CUDA graph replay (default); eager: per-kernel launches")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    _quiet_stdout()
    if args.impl == "reference":
        run_reference(args)
        return
    args.warmup = max(args.warmup, 3)
    line = run_ours(args)
    if line is None:
        return
    if line["n_gpus"] == 1 and not args.no_cpu_baseline:
        sec, cores, sample = cpu_step_seconds(1, warmup=0)
        line["cpu_baseline"] = {"value": 1.0 / sec, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample}
    else:
        line["cpu_baseline"] = None
    _emit(line)


if __name__ == "__main__":
    main()
