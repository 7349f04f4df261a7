#!/usr/bin/env python
"""bench.py — DN-Splatter train step throughput on B200 (BASELINE.json metric).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--config cfg4|cfg2]

Default workload (config.workload starts with "cfg4" = BASELINE.json configs[3], the configuration the north-star
target is quoted on): 1M Gaussians, 1920x1080, 8 training views, one camera view per GPU per iteration (at --gpus 8
the step's camera batch of 8 is sharded over the 8 GPUs): get_outputs (RGB+ED rasterization + legacy normals pass)
-> losses -> backward -> [gradient all-reduce] -> Adam -> after_train, one CUDA-graph replay per iteration.  The
FusionSense-sized scene (cfg2 = configs[1]: 300k Gaussians, 640x480, 9 views) rides along in the same line as
`secondary` (value, e2e and, at N = 1, the shim-only eager value).  One JSON line on stdout (rank 0); DESIGN.md
§Measurement defines every key.

--impl reference: a real gsplat CUDA build if one is importable from baseline/_ref (reported as
`gsplat_cuda_baseline`; never the case so far: gsplat==1.0.0 is not vendored and there is no wheel), else the same
step through the CPU oracle (oracle/gsplat_ref.py standing in for gsplat) on the box's host cores, each step a
bounded sample of the workload.
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

CONFIGS = {
    # BASELINE.json configs[1]
    "cfg2": dict(n=300_000, w=640, h=480, views=9, kind="bunny", cfg_id=2,
                 workload="cfg2: DN-Splatter train step, 9 views 640x480 synthetic bunny, 300k Gaussians, "
                          "1 view/GPU/iter"),
    # BASELINE.json configs[3]: camera batch 8 sharded across 8 GPUs = one view per GPU per iteration
    "cfg4": dict(n=1_000_000, w=1920, h=1080, views=8, kind="random", cfg_id=4,
                 workload="cfg4: DN-Splatter train step (render + backprop + Adam), 1M Gaussians 1920x1080, 8 views, "
                          "1 view/GPU/iter (the camera batch of 8 is sharded over the GPUs at N = 8)"),
    # BASELINE.json configs[4]: camera batch 32 per iteration dealt over the GPUs (32 / N views per rank per iteration,
    # rendered one after the other inside the one captured step, gradients accumulated, one Adam step), 8 distinct views
    "cfg5": dict(n=3_000_000, w=3840, h=2160, views=8, kind="random", cfg_id=5, global_views=32,
                 workload="cfg5: DN-Splatter train step, 3M Gaussians 3840x2160, camera batch 32 per iteration sharded "
                          "over the GPUs (32 / N views per rank, 8 distinct views cycled), one Adam step per iteration"),
}
METRIC = "dn_splatter_train_iter_per_s"
UNIT = "iter/s"


def _peaks():
    p = ROOT / "MEASURED_PEAKS.json"
    if p.exists():
        d = json.loads(p.read_text())
        return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def _ncu_traffic(cfg_name: str):
    """DRAM bytes (dram__bytes_read.sum + dram__bytes_write.sum) of ONE launch of the roofline kernel on this
    workload (+ its source text and the issue-slot / pipe figures of the same launch), from the committed
    `ncu --set full` capture (profiles/raster_bwd_traffic_<cfg>.json, written by
    tools/ncu_traffic.py from the raw page); None when no capture is committed for this configuration."""
    for name in (f"raster_bwd_traffic_{cfg_name}.json",) + (("raster_bwd_traffic.json",) if cfg_name == "cfg2" else ()):
        p = ROOT / "profiles" / name
        if p.exists():
            d = json.loads(p.read_text())
            return d.get("dram_bytes_per_launch"), d.get("source"), d.get("issue")
    return None, None, None


def _fp32_roofline(pairs, ktimes, D, clocks, key=None):
    """Compositing is bound by the FP32 pipe, so next to the HBM fraction BASELINE.json asks for: algorithmic
    pair-flops (forward Q (14 + 2 D), backward Q (40 + 6 D), Q = blended (pixel, entry) pairs counted
    on the device by fsb_raster_pair_count) over the kernels' CUDA-event times, against 148 SMs x 128 FP32 lanes x
    2 flop x the SM clock sampled under load."""
    key = key or f"D{D}"
    q = pairs.get("D4") if key == "D4+3" else pairs.get(f"D{D}")
    if not q:
        return None
    mhz = (clocks or {}).get("sm_max_mhz") or 1965.0
    peak = 148 * 128 * 2 * mhz * 1e6 / 1e12
    out = {"pairs_blended": q["blended"], "pairs_visited": q["visited"], "peak_tflops": peak,
           "peak_source": f"148 SM x 128 lanes x 2 x {mhz:.0f} MHz (nominal FP32 FMA rate, not measured)"}
    for name, flop_per_pair in (("raster_fwd", 14 + 2 * D), ("raster_bwd", 40 + 6 * D)):
        ms = ktimes.get(f"{name}_{key}", (float("nan"), 0))[0]
        if ms == ms and ms > 0:
            tf = q["blended"] * flop_per_pair / (ms * 1e-3) / 1e12
            out[name] = {"flop_per_pair": flop_per_pair, "achieved_tflops": tf, "frac": tf / peak, "kernel_ms": ms}
    return out


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 500 ms from before the warm-up on; every sample is stamped on
    arrival and `window(t0, t1)` reports the ones that fall inside a timed region (the neighbouring ones, flagged, if
    the region was too short to catch one).  The period is deliberately long: every nvidia-smi query stalls the GPU
    for ~3.6 ms (r01l: five samples inside a 219 ms region cost 8 % of the measured rate)."""

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

    def window(self, t0=None, t1=None):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        rows = list(self.rows)
        have = t0 is not None and t1 is not None
        inside = [r for t, r in rows if have and t0 <= t <= t1 + 0.05]
        # a region shorter than the sampling period: the samples next to it (warm-up steps before, e2e leg after —
        # the same workload) stand in
        near = [r for t, r in rows if have and t0 - 0.6 <= t <= t1 + 0.6]
        window = "timed region" if inside else ("timed region +- 0.6 s (warm-up / e2e leg)" if near else "whole run")
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
                "reasons": sorted(reasons), "samples": len(sm), "window": window}

    def stop(self):
        if self.proc is not None:
            time.sleep(0.1)
            self.proc.terminate()


def build_model(cfg_name, device, gsplat_module=None, fused=True, crop=None):
    """`crop` = (x0, y0, w, h): the same scene seen through a sub-window of the image (principal point shifted), the
    bounded sample of the CPU legs."""
    from fusionsense_b200.dn_step import DNSplatterStep, DNSplatterStepConfig
    from fusionsense_b200.synthetic import make_scene

    c = CONFIGS[cfg_name]
    scene = make_scene(c["n"], c["w"], c["h"], n_views=c["views"], cfg_id=c["cfg_id"], kind=c["kind"])
    if crop is not None:
        x0, y0, w, h = crop
        scene.Ks = scene.Ks.clone()
        scene.Ks[:, 0, 2] -= x0
        scene.Ks[:, 1, 2] -= y0
        scene.width, scene.height = w, h
    cfg = DNSplatterStepConfig(fused_optimizer=fused)
    torch_losses = None
    if not fused:
        cfg.stop_split_at = 0  # the CPU oracle has no absgrad side channel; after_train is skipped there
        from oracle import dn_losses_ref as torch_losses  # CPU legs only: the product package has no torch losses
    return DNSplatterStep(scene, cfg, device=device, step=3000, gsplat_module=gsplat_module, torch_losses=torch_losses)


# ---------------------------------------------------------------------------------------------
# reference arm: gsplat CUDA build when importable, else the CPU oracle
# ---------------------------------------------------------------------------------------------
def _cpu_step_once(cfg_name, crop, steps, warmup):
    import torch
    from oracle import gsplat_ref as ref

    model = build_model(cfg_name, "cpu", gsplat_module=ref, fused=False, crop=crop)
    targets = model.render_targets(0)
    ts = []
    for i in range(warmup + steps):
        t0 = time.perf_counter()
        model.train_iteration(0, targets)
        dt = time.perf_counter() - t0
        if i >= warmup:
            ts.append(dt)
    del model
    return statistics.median(ts)


def cpu_step_seconds(cfg_name: str, steps: int, warmup: int):
    """Time the oracle-driven train step on the host cores -> (seconds per FULL step, cores, sample text).

    cfg2 runs whole steps (~6 s each).  cfg4 (1M Gaussians, 1080p) would take minutes per step, so its sample is
    bounded: the same 1M-Gaussian scene rendered and differentiated through small sub-windows spread over the image
    (every Gaussian is still projected, binned and culled), see below."""
    import torch

    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    c = CONFIGS[cfg_name]
    if cfg_name == "cfg2":
        sec = _cpu_step_once(cfg_name, None, steps, warmup)
        sample = (f"{steps} full train step(s) (RGB+ED pass, normals pass, losses, backward, torch Adam) of the same "
                  f"300k-Gaussian 640x480 scene through oracle/gsplat_ref.py, torch CPU fp32, {cores} threads")
        return sec, cores, sample
    # Stratified sample: the image is cut into a 3 x 2 grid of cells and the step is run through a 160 x 96 window in
    # the middle of every cell (all 1M Gaussians are projected, binned and culled each time), plus once through a
    # 16 x 16 corner window whose time is the per-Gaussian part t0 (projection, SH, Adam).  Full step =
    # t0 + sum_cells (t_cell - t0) * cell area / window area.  (Two centred windows and a linear fit, the first version,
    # extrapolated the densest part of the frame to all of it: 248 s on the box where this form gives less.)
    W, H = c["w"], c["h"]
    gx, gy, ww, wh = 3, 2, 160, 96
    t0 = _cpu_step_once(cfg_name, (0, 0, 16, 16), 1, 0)
    cell_w, cell_h = W // gx, H // gy
    times, per_px = [], 0.0
    for j in range(gy):
        for i in range(gx):
            x0 = i * cell_w + (cell_w - ww) // 2
            y0 = j * cell_h + (cell_h - wh) // 2
            t = _cpu_step_once(cfg_name, (x0 // 16 * 16, y0 // 16 * 16, ww, wh), 1, 0)
            times.append(t)
            per_px += max(0.0, t - t0) / (ww * wh) * (cell_w * cell_h)
    sec = t0 + per_px
    sample = (f"{gx * gy + 1} train steps of the same {c['n']}-Gaussian scene, all Gaussians projected each time: one "
              f"through a 16x16 corner window ({t0:.1f} s = the per-Gaussian part) and one through a {ww}x{wh} window in "
              f"the middle of every cell of a {gx}x{gy} grid over the {W}x{H} image ("
              + ", ".join(f"{t:.1f}" for t in times) + f" s); full step = {t0:.1f} s + sum over cells of (t - {t0:.1f} s) "
              f"x cell area / window area = {sec:.1f} s; oracle/gsplat_ref.py, torch CPU fp32, {cores} threads")
    return sec, cores, sample


def _real_gsplat():
    """A gsplat that is NOT this repository's shim (the driver's baseline/_ref install), or None."""
    ref_dir = ROOT / "baseline" / "_ref"
    if not ref_dir.is_dir():
        return None
    sys.path.insert(0, str(ref_dir))
    try:
        import importlib

        g = importlib.import_module("gsplat")
        f = Path(getattr(g, "__file__", "") or "").resolve()
        if str(f).startswith(str(ROOT / "fusionsense_b200")) or str(f).startswith(str(ROOT / "shim")):
            return None
        from gsplat.rendering import rasterization  # noqa: F401
        return g
    except Exception:  # noqa: BLE001
        return None
    finally:
        sys.path.remove(str(ref_dir))


def gsplat_cuda_step_ms(g, cfg_name, steps, warmup):
    """The two gsplat calls of dn_model.py:570-591 / :644-653 (rasterization RGB+ED with absgrad, then the legacy
    rasterize_gaussians normals pass) forward + backward on the real gsplat CUDA build, same scene, CUDA events."""
    import torch
    from fusionsense_b200.synthetic import make_scene

    c = CONFIGS[cfg_name]
    sc = make_scene(c["n"], c["w"], c["h"], n_views=c["views"], cfg_id=c["cfg_id"], kind=c["kind"]).to("cuda")
    P = lambda t: t.clone().requires_grad_(True)  # noqa: E731
    means, quats, scales, opac = P(sc.means), P(sc.quats), P(sc.scales), P(sc.opacities)
    dc, rest = P(sc.features_dc), P(sc.features_rest)
    normals = torch.nn.functional.normalize(torch.randn_like(sc.means), dim=-1).requires_grad_(True)

    def step(v):
        colors = torch.cat((dc[:, None, :], rest), dim=1)
        render, alpha, info = g.rendering.rasterization(
            means=means, quats=quats / quats.norm(dim=-1, keepdim=True), scales=torch.exp(scales),
            opacities=torch.sigmoid(opac).squeeze(-1), colors=colors, viewmats=sc.viewmats[v:v + 1],
            Ks=sc.Ks[v:v + 1], width=c["w"], height=c["h"], tile_size=16, packed=False, near_plane=0.01,
            far_plane=1e10, render_mode="RGB+ED", sh_degree=3, sparse_grad=False, absgrad=True,
            rasterize_mode="classic")
        nim = g.rasterize_gaussians(info["means2d"][0].detach(), info["depths"][0], info["radii"][0],
                                    info["conics"][0], info["tiles_per_gauss"][0], normals, torch.sigmoid(opac),
                                    c["h"], c["w"], 16)
        (render.sum() + nim.sum()).backward()

    for i in range(warmup):
        step(i % c["views"])
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for i in range(steps):
        step(i % c["views"])
    e.record()
    torch.cuda.synchronize()
    return s.elapsed_time(e) / steps


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    c = CONFIGS[args.config]
    steps = max(1, min(args.steps, 3))
    gsplat_cuda = None
    g = _real_gsplat()
    if g is not None:
        try:
            import torch

            if torch.cuda.is_available():
                ms = gsplat_cuda_step_ms(g, args.config, max(3, min(args.steps, 20)), max(1, min(args.warmup, 5)))
                gsplat_cuda = {"ms_render_fwd_bwd": ms, "unit": "ms", "version": getattr(g, "__version__", "?"),
                               "what": "gsplat.rendering.rasterization (RGB+ED, absgrad) + gsplat.rasterize_gaussians "
                                       "forward + backward on the same scene, CUDA events"}
        except Exception as exc:  # noqa: BLE001
            gsplat_cuda = {"error": f"{type(exc).__name__}: {exc}"[:300]}
    sec, cores, sample = cpu_step_seconds(args.config, steps, warmup=min(args.warmup, 1))
    v = 1.0 / sec
    line = {
        "impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": steps,
        "warmup": min(args.warmup, 1), "ms_per_step": sec * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": c["workload"], "gaussians": c["n"], "width": c["w"], "height": c["h"],
                   "note": "gsplat==1.0.0 is not vendored in the reference tree and not installed; the reference's "
                           "CPU path is its algorithm restated in oracle/ (kind=port)"},
        "cpu_baseline": {"value": v, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gsplat_cuda_baseline": gsplat_cuda if gsplat_cuda is not None else
        "unavailable: no gsplat other than this repository's shim is importable (baseline/_ref absent)",
        "gpu_launches": 0,
    }
    _emit(line)


# ---------------------------------------------------------------------------------------------
# our arm
# ---------------------------------------------------------------------------------------------
_FULL_AFFINITY = None  # the affinity the process started with (the CPU-baseline leg runs on all of it)


def _bind_to_gpu_numa_node(local_rank: int):
    """Pin this rank's threads to the CPUs NVML names as closest to its GPU, BEFORE any pinned host buffer is
    allocated: the end-to-end leg copies 58 MB of targets per step from pinned memory, and with 8 ranks on a two-socket
    host half of those copies otherwise cross the socket interconnect (r02m, 8 GPUs: e2e 1200 views/s against a
    device-resident 1980).  Returns the number of CPUs in the mask, or None when NVML cannot say."""
    global _FULL_AFFINITY
    if os.environ.get("FSB_BIND_NUMA", "1") == "0":
        return None
    try:
        _FULL_AFFINITY = os.sched_getaffinity(0)
        import pynvml

        pynvml.nvmlInit()
        visible = os.environ.get("CUDA_VISIBLE_DEVICES")
        index = int(visible.split(",")[local_rank]) if visible and visible.split(",")[local_rank].isdigit() else local_rank
        handle = pynvml.nvmlDeviceGetHandleByIndex(index)
        n_words = (os.cpu_count() + 63) // 64
        mask = pynvml.nvmlDeviceGetCpuAffinity(handle, n_words)
        cpus = {64 * w + b for w, word in enumerate(mask) for b in range(64) if (int(word) >> b) & 1}
        cpus &= os.sched_getaffinity(0)
        if cpus:
            os.sched_setaffinity(0, cpus)
            return len(cpus)
    except Exception:  # noqa: BLE001 - no NVML, a container without the topology: leave the affinity alone
        return None
    return None


class Ctx:
    """Process-wide state of one bench run (one rank)."""

    def __init__(self):
        import torch
        import torch.distributed as dist

        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.rank = int(os.environ.get("RANK", "0"))
        self.local_rank = int(os.environ.get("LOCAL_RANK", "0"))
        if not torch.cuda.is_available():
            raise SystemExit("bench.py (impl=ours) needs a CUDA device: fusionsense_b200 has no CPU fallback")
        torch.cuda.set_device(self.local_rank)
        self.device = torch.device("cuda", self.local_rank)
        self.cpu_affinity = _bind_to_gpu_numa_node(self.local_rank)
        if self.world > 1:
            dist.init_process_group("nccl", device_id=self.device)
        # N > 1 gradient exchange: "peer" = this library's kernels over NVLink peer memory (default), "nccl" = one flat
        # NCCL all-reduce; both captured in the step's graph
        self.exchange = os.environ.get("FSB_EXCHANGE", "peer")
        self.sampler = ClockSampler(self.local_rank)
        if self.rank == 0:
            self.sampler.start()


def run_workload(ctx: Ctx, cfg_name: str, steps: int, warmup: int, mode: str, detail: bool, literal_leg: bool,
                 refine_leg: bool = False):
    """Time one configuration: resident-targets leg, end-to-end leg and (detail) the per-kernel roofline legs.
    Returns a dict on every rank (timings are max-over-ranks)."""
    import torch
    import torch.distributed as dist

    from fusionsense_b200 import ops
    from fusionsense_b200._abi import lib
    from fusionsense_b200.dist import GradSync

    world, rank, device = ctx.world, ctx.rank, ctx.device
    c = CONFIGS[cfg_name]
    n_views = c["views"]
    vpi = max(1, c.get("global_views", world) // world)  # views per rank per iteration
    model = build_model(cfg_name, device)
    views = list(range(n_views))
    # RGB and normal targets hold 8-bit values, as the PNGs of a FusionSense dataset do (images/rgb_i.png,
    # normals_from_pretrain/*.png); the host side of the e2e leg keeps them as uint8 (nerfstudio caches uint8 images)
    # and the float32 `x / 255.0` of splatfacto's get_gt_img / dn_dataset.py:205 happens on the device after the copy
    # (fsb_u8_to_unit_float, the reference's bits): 20.7 MB per 1080p view cross PCIe instead of 58 MB.  Depth is
    # float32.  The resident leg trains on the very same float32 values (graph_step.eight_bit_targets).
    from fusionsense_b200.compose import u8_to_unit_float
    from fusionsense_b200.graph_step import eight_bit_targets

    dev_targets, host_targets = {}, {}
    for v in views:
        dev_targets[v], host_targets[v] = eight_bit_targets(model.render_targets(v))
    if os.environ.get("FSB_U8_TARGETS", "1") == "0":  # A/B: the same values cross PCIe as float32 (58 MB per 1080p view)
        host_targets = {v: {k: t.cpu().pin_memory() for k, t in d.items()} for v, d in dev_targets.items()}
    h2d_bytes = sum(t.numel() * t.element_size() for t in host_targets[0].values()) * min(vpi, n_views)
    params = [model.gauss_params[k] for k in model.config.lrs]
    graph_mode = mode == "graph"

    # one flat NCCL all-reduce (SUM) over all Gaussian parameter gradients (59 floats per Gaussian) plus the
    # overflow flag of the static-capacity step; the loss is pre-scaled by 1/world, so the sum is the mean
    allreduce_grads = GradSync()
    runner = None
    if graph_mode:
        from fusionsense_b200.graph_step import GraphedDNSplatterStep

        sync = None
        if world > 1:
            sync = allreduce_grads
            if ctx.exchange == "peer":
                # our own kernels over NVLink peer memory (csrc/grad_exchange.cu) instead of the NCCL all-reduce
                from fusionsense_b200.dist import PeerGradExchange

                try:
                    sync = PeerGradExchange()
                    sync.probe(device)
                except Exception as exc:  # noqa: BLE001  (no symmetric-memory support on this box: say so, use NCCL)
                    print(f"bench: PeerGradExchange unavailable ({type(exc).__name__}: {exc}); using NCCL",
                          file=sys.stderr)
                    ctx.exchange = "nccl (peer exchange unavailable)"
                    sync = allreduce_grads
        runner = GraphedDNSplatterStep(model, dev_targets, grad_sync=sync, loss_scale=1.0 / world, views_per_iter=vpi)

    def eager_step(m, i, batch):
        v = (i * world + rank) % n_views
        for opt in m.optimizers.values():
            opt.zero_grad(set_to_none=True)
        outputs = m.get_outputs(v)
        loss = m.get_loss_dict(outputs, batch(v))["main_loss"]
        (loss / world if world > 1 else loss).backward()
        if world > 1:
            allreduce_grads([m.gauss_params[k] for k in m.config.lrs])
        m.optimizers["means"].param_groups[0]["lr"] = m._means_lr()
        m.optimizer_step()
        m.after_train()
        m.step += 1
        return loss

    def resident(v):
        return dev_targets[v]

    def from_host(v):
        out = {k: t.to(device, non_blocking=True) for k, t in host_targets[v].items()}
        return {k: u8_to_unit_float(t, recip=(k == "image")) if t.dtype == torch.uint8 else t for k, t in out.items()}

    def one_step(i, staged, read_loss, n_total=1 << 30):
        """One training iteration; `staged`: this step's targets come from pinned host memory; `read_loss`: the
        step's result is read back to the host."""
        if runner is None:
            loss = eager_step(model, i, from_host if staged else resident)
            if read_loss:
                float(loss)
            return
        vs = [((i * vpi + j) * world + rank) % n_views for j in range(vpi)]
        if staged and i == 0:
            for v in dict.fromkeys(vs):
                runner.stage_async(v, host_targets[v])
        runner.train_iteration(vs if vpi > 1 else vs[0])  # waits for its views' copies
        if staged and i + 1 < n_total:
            # the next step's inputs travel while this step's kernels run (copy stream, pinned source); a view that
            # the running step reads as well is re-staged after it (stage_async orders the copy behind the reader)
            for vn in dict.fromkeys(((i + 1) * vpi + j) * world + rank for j in range(vpi)):
                runner.stage_async(vn % n_views, host_targets[vn % n_views])
        if read_loss:
            # 32-byte D2H read of [loss, overflow count, n_isects x2] per step; the host consumes step i - 1's
            # values here and the last step's after the loop (timed() drains the ring before the closing event)
            runner.read_result_async()

    def timed(n_steps, step_fn, drain=None):
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        for i in range(n_steps):
            step_fn(i)
        if drain is not None:
            drain()
        e.record()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        ms = torch.tensor([s.elapsed_time(e)], device=device)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms.item())

    # warm-up (also primes the caching allocator and, in graph mode, captures the step).  The W requested steps, then
    # more of the same until the GPU has been busy for 0.4 s: the first replays after a capture run at ramping clocks
    # (r02d: the cfg2 leg measured 843 iter/s right after 5 warm-up steps and 1121 iter/s one leg later)
    for i in range(warmup):
        one_step(i, False, False)
    torch.cuda.synchronize()
    t_w, warm_extra = time.perf_counter(), 0
    while True:
        go = torch.tensor([1 if (time.perf_counter() - t_w < 0.4 and warm_extra < 400) else 0], device=device)
        if world > 1:
            dist.broadcast(go, src=0)  # the steps carry collectives: every rank must run the same number
        if not int(go.item()):
            break
        for i in range(5):
            one_step(warmup + warm_extra + i, False, False)
        warm_extra += 5
        torch.cuda.synchronize()
    if runner is not None and runner.poll()["new_overflows"]:
        for i in range(warmup):  # capacity grew: capture again before timing
            one_step(i, False, False)
        torch.cuda.synchronize()

    # Nothing the capture froze may change inside a timed leg: the binary-opacity window of dn_model.py:497-503 opens at
    # step % 3000 == 201 (r02i: starting at step 3000, the 115 warm-up + 200 timed steps crossed 3201 and the re-capture
    # landed in the timed region: 4.83 ms per step against 3.67 ms in the next leg).  Every timed leg starts at 3300.
    BENCH_STEP = 3300
    model.step = BENCH_STEP
    one_step(0, False, False)  # re-capture for the new step count, outside the timed region
    torch.cuda.synchronize()
    model.step = BENCH_STEP
    launches0 = lib.fsb_launch_count()
    replays0 = runner.replays if runner else 0
    captures0 = runner.captures if runner is not None else 0
    profiling = os.environ.get("FSB_PROFILE") == cfg_name  # ncu --profile-from-start off: capture the timed steps only
    if profiling:
        torch.cuda.profiler.start()
    t_region0 = time.perf_counter()
    ms_total = timed(steps, lambda i: one_step(i, False, False, steps))
    t_region1 = time.perf_counter()
    if profiling:
        torch.cuda.profiler.stop()
    launches = lib.fsb_launch_count() - launches0
    captures_in_value_leg = (runner.captures - captures0) if runner is not None else 0
    graph_info = None
    if runner is not None:
        info = runner.poll()
        if info["new_overflows"]:
            raise SystemExit(f"bench: {info['new_overflows']} timed step(s) overflowed the intersection capacity; rerun")
        # replays launch the captured kernels without passing through the library's entry points
        launches += (runner.replays - replays0) * runner.launches_per_replay
        graph_info = {"captures": runner.captures, "capacity": runner.capacity, "n_isects": info["n_isects"],
                      "n_isects_normals": info["n_isects_normals"],
                      "libfsb200_launches_per_replay": runner.launches_per_replay,
                      "graph_launches_per_step": 1 if runner.graph_tail is None else 2}
    clocks = ctx.sampler.window(t_region0, t_region1) if rank == 0 else None
    captures_before_e2e = runner.captures if runner is not None else 0
    # the staged path has first-use costs of its own (copy stream, staging buffers, the pinned result ring's
    # cudaHostAlloc, the conversion kernel's lazy load: tens to hundreds of ms on some boxes — r02s measured the 200-step
    # cfg2 e2e leg anywhere between 366 and 1166 it/s because of them): warm it up like the resident leg
    for i in range(max(3, min(warmup, 10))):
        one_step(i, True, True, 1 << 30)
    if runner is not None:
        runner.poll()
    torch.cuda.synchronize()
    model.step = BENCH_STEP
    ms_e2e = timed(steps, lambda i: one_step(i, True, True, steps),
                   drain=(lambda: runner.poll()) if runner is not None else None)
    if runner is not None and runner.poll()["overflowed_steps"]:
        raise SystemExit("bench: a step of the e2e leg overflowed the intersection capacity; rerun")
    if runner is not None:
        graph_info["captures_inside_timed_legs"] = (runner.captures - captures_before_e2e) + captures_in_value_leg

    ms_per_step = ms_total / steps
    out = {
        "cfg": cfg_name, "value": world * vpi * 1e3 / ms_per_step, "ms_per_step": ms_per_step,
        "optimizer_steps_per_s": 1e3 / ms_per_step, "views_per_s": world * vpi * 1e3 / ms_per_step,
        "views_per_iter_per_gpu": vpi,
        "e2e": {"value": world * vpi * steps * 1e3 / ms_e2e, "unit": UNIT, "h2d_bytes_per_step": h2d_bytes,
                "d2h_bytes_per_step": 32 if graph_mode else 4},
        "gpu_launches": int(launches), "gpu_launches_per_step": launches / steps, "clocks": clocks,
        "graph": graph_info, "warmup_actual": warmup + warm_extra, "fused_outputs": bool(model.config.fused_outputs),
        "prune_lists": bool(model.config.prune_lists), "step_metrics": bool(model.config.step_metrics),
        "host_targets": {k: str(t.dtype).replace("torch.", "") for k, t in host_targets[0].items()},
    }

    if detail:
        # per-kernel CUDA-event times of the same workload: eager launches (a replayed graph offers no place to
        # record events between its kernels), same kernels, same sizes, same stream, after the timed legs
        with ops.kernel_timer.collect(pad_cycles=400_000):
            for i in range(6):
                eager_step(model, i, resident)
            ktimes = ops.kernel_timer.summary()
        # pair counts (Q of SURVEY.md §8d) of one more eager step, counted by fsb_raster_pair_count after each forward
        ops.pair_probe.enabled = True
        eager_step(model, 5, resident)
        torch.cuda.synchronize()
        ops.pair_probe.enabled = False
        pairs = ops.pair_probe.summary()
        out["roofline"] = _roofline(cfg_name, model, params, ktimes, pairs, clocks)

    if refine_leg and runner is not None:
        # densify / prune inside a timed run (dn_model.py:326-451 every refine_every = 100 iterations):
        # 300 iterations from a step count at which the refine steps really densify, the (synchronised) refinement,
        # the re-capture on the new Gaussian count and the capacity re-probe all inside the timed region
        from fusionsense_b200 import dist as fdist

        model.step = 3100
        runner.graph = None
        n_before = model.num_points
        cap0, refines = runner.captures, []

        host_s = {"refinement": 0.0}
        cap_s0 = runner.capture_seconds

        def step_with_refine(i):
            one_step(i, False, False)
            if model.step % model.config.refine_every == 0:
                t0 = time.perf_counter()
                runner.poll()
                if world > 1:
                    fdist.synchronised_refinement(model, model.optimizers, model.step, seed=7)
                else:
                    model.refinement_after()
                refines.append(model.num_points)
                host_s["refinement"] += time.perf_counter() - t0

        n_ref = 300
        ms_ref = timed(n_ref, step_with_refine, drain=lambda: runner.poll())
        out["with_refinement"] = {"value": world * n_ref * 1e3 / ms_ref, "unit": UNIT, "steps": n_ref,
                                  "ms_per_step": ms_ref / n_ref, "refine_every": model.config.refine_every,
                                  "gaussians_before": n_before, "gaussians_after_each_refine": refines,
                                  "recaptures": runner.captures - cap0,
                                  # host wall time inside the timed region that is not graph replay: where a slow leg
                                  # spent its time (poll + refinement_after incl. its device sync; warm-up + capture)
                                  "host_seconds": {"refinement": round(host_s["refinement"], 4),
                                                   "capture": round(runner.capture_seconds - cap_s0, 4)},
                                  "what": "300 captured iterations with refinement_after every 100 (densify + cull + "
                                          "Adam-state rebuild, statistics all-reduced and a shared split RNG at N > 1), "
                                          "re-captures included"}
    if literal_leg:
        out["shim_only_eager"] = _reference_model_leg(ctx, cfg_name, model, dev_targets, min(steps, 100), warmup, timed)
    del model, dev_targets, host_targets
    torch.cuda.empty_cache()
    return out


def _reference_model_leg(ctx, cfg_name, model, dev_targets, n_steps, warmup, timed):
    """What an UNMODIFIED reference gets from the drop-in alone: the reference's own dn_splatter/dn_model.py
    (baseline/_ref, the sanctioned --no-deps install) driven the way nerfstudio's Trainer does — get_outputs with its
    torch glue around our rasterization() / rasterize_gaussians(), its torch loss classes, one torch.optim.Adam per
    group, splatfacto's after_train — eager launches, on stub nerfstudio / torchmetrics packages (tests/stubs: neither
    is installed here).  None when baseline/_ref is absent."""
    try:
        from tests import stubs
        from tests.stubs import harness
    except ImportError:
        return None
    if not stubs.reference_available():
        return {"unavailable": "baseline/_ref/dn_splatter is not installed on this box"}
    scene = harness.gl_scene(model.scene.to("cpu"))
    _, ref = harness.build_reference_model(scene, 3001, ctx.device)
    opts = harness.build_optimizers(ref, dict(model.config.lrs))
    n_views = scene.viewmats.shape[0]
    cams = [harness.camera_for(scene, v, ctx.device) for v in range(n_views)]
    state = {"step": 3001}

    def one(i):
        v = (i * ctx.world + ctx.rank) % n_views
        if state["step"] % 100 == 0:
            state["step"] += 1  # dn_model.py:905 writes a jpg through matplotlib every 100 steps
        harness.train_iteration(ref, opts, cams[v], dev_targets[v], state["step"])
        state["step"] += 1

    for i in range(max(3, warmup)):
        one(i)
    ms = timed(n_steps, one)
    return {"value": ctx.world * n_steps * 1e3 / ms, "unit": UNIT, "steps": n_steps,
            "what": "the reference's own DNSplatterModel (baseline/_ref/dn_splatter/dn_model.py, unmodified) on this "
                    "repository's gsplat drop-in: its torch glue, torch losses, torch.optim.Adam per group, eager"}


def _roofline(cfg_name, model, params, ktimes, pairs, clocks):
    """Roofline of the dominant kernel (raster backward of the RGB+ED pass, D = 4) and the per-stage table: algorithmic
    bytes (SURVEY.md §8d formulas with this step's N, Nv, I, P, K) over the stage's mean CUDA-event time."""
    from fusionsense_b200.gsplat.cuda_legacy import _wrapper as _legacy

    c = CONFIGS[cfg_name]
    I = int(_legacy._LAST_BINNING.get("n_isects", 0))  # intersections of the last eager step
    Nv = int((model.radii > 0).sum().item())
    P = c["w"] * c["h"]
    # fused_passes (default): ONE kernel composites RGB + depth (4 channels) and the normals (3): D = 7 in the
    # SURVEY §8d formulas, one read of the list, one write of the per-Gaussian gradients
    fused = "raster_bwd_D4+3" in ktimes
    D = 7 if fused else 4
    key = "D4+3" if fused else "D4"
    bwd_ms = ktimes.get(f"raster_bwd_{key}", (float("nan"), 0))[0]
    bwd_bytes = I * (28 + 4 * D) + P * (4 * D + 12) + Nv * (32 + 4 * D)
    peak, peak_src = _peaks()
    achieved = bwd_bytes / (bwd_ms * 1e-3) / 1e9 if bwd_ms == bwd_ms and bwd_ms > 0 else None
    traffic, traffic_src, issue = _ncu_traffic(cfg_name)
    N, K, C = c["n"], 16, 1
    tiles = ((c["w"] + 15) // 16) * ((c["h"] + 15) // 16)
    key_bytes = -(-(32 + max(1, (tiles - 1).bit_length())) // 8)
    from fusionsense_b200 import ops as _ops

    two_level = bool(model.config.prune_lists) and _ops.TWO_LEVEL_BINNING
    # two-level binning: C N (key 8 B + value 4 B) pairs sorted on 32 depth bits (histogram read + 4 passes of read +
    # write), then the I 8-byte entries on the tile bits (histogram read + 2 passes + the 4-byte flatten ids)
    sort_bytes = (C * N * (8 + 4 * 24) + I * (8 + 2 * 16 + 4)) if two_level else (I * 8 + I * 24 * key_bytes)
    stage_bytes = {
        "project_sh_fwd": C * N * 68 + Nv * (12 * K + 12),
        "project_sh_bwd": C * N * 4 + Nv * (76 + 12 * K) + N * (40 + 12 * K),
        "radix_sort": sort_bytes,
        "adam_multi": sum(p.numel() for p in params) * 28,
        "raster_fwd_D4": I * (28 + 16) + P * (16 + 8), "raster_fwd_D3": I * (28 + 12) + P * (12 + 8),
        "raster_bwd_D4": I * (28 + 16) + P * (16 + 12) + Nv * (32 + 16),
        "raster_bwd_D3": I * (28 + 12) + P * (12 + 12) + Nv * (32 + 12),
        "raster_fwd_D4+3": I * (28 + 28) + P * (28 + 8), "raster_bwd_D4+3": I * (28 + 28) + P * (28 + 12) + Nv * (32 + 28),
    }
    stages = {}
    for name, nbytes in stage_bytes.items():
        ms = ktimes.get(name, (float("nan"), 0))[0]
        if ms == ms and ms > 0:
            gbs = nbytes / (ms * 1e-3) / 1e9
            stages[name] = {"algorithmic_bytes": int(nbytes), "ms": round(ms, 4), "gbs": round(gbs, 1),
                            "hbm_frac": round(gbs / peak, 4),
                            "bound": "fp32" if name.startswith("raster") else "hbm"}
    fwd_ms = ktimes.get(f"raster_fwd_{key}", (0, 0))[0]
    return {
        "kernel": (("raster_bwd2_kernel<7,4,2> (two pixels per lane; " if tiles * C >= 4 * 148 * 4 else
                    "raster_bwd_kernel<7,4,2> (") + "RGB + expected depth + normals in one walk)" if fused
                   else "raster_bwd_kernel<4,4,2> (RGB+ED pass)"), "bound": "hbm", "achieved": achieved, "peak": peak,
        "unit": "GB/s", "frac": (achieved / peak) if achieved else None, "traffic": traffic,
        "traffic_source": traffic_src, "peak_source": peak_src, "algorithmic_bytes": bwd_bytes, "kernel_ms": bwd_ms,
        # the resource that does bind this kernel (issue slots / pipes of the same ncu launch the traffic comes from)
        "issue": issue,
        "n_isects": I, "n_visible": Nv,
        "note": "compositing is FP32/MUFU-bound, not HBM-bound (SURVEY.md §8d); the HBM fraction is reported as "
                "BASELINE.json asks, the pipe utilisation is in profiles/",
        "fp32": _fp32_roofline(pairs, ktimes, D, clocks, key),
        "kernel_ms_all": {k: round(v[0], 4) for k, v in sorted(ktimes.items())},
        "stages": stages,
        "raster_fwd_mpix_per_s": (P / (fwd_ms * 1e-3) / 1e6) if fwd_ms > 0 else None,
        "raster_fwd_bwd_mpix_per_s": (P / ((fwd_ms + bwd_ms) * 1e-3) / 1e6) if bwd_ms == bwd_ms and bwd_ms > 0 else None,
    }


def run_ours(args):
    import torch
    import torch.distributed as dist

    ctx = Ctx()
    world, rank = ctx.world, ctx.rank
    graph_mode = args.mode == "graph"
    main = run_workload(ctx, args.config, args.steps, args.warmup, args.mode, detail=True, literal_leg=False)
    secondary = None
    if args.config != "cfg2" and not args.no_secondary:
        # the FusionSense-sized scene: long enough a region for stable numbers (1 ms steps)
        secondary = run_workload(ctx, "cfg2", max(args.steps, 200), args.warmup, args.mode, detail=False,
                                 literal_leg=(world == 1), refine_leg=True)
    ctx.sampler.stop()
    line = None
    if rank == 0:
        c = CONFIGS[args.config]
        line = {
            "metric": METRIC, "value": main["value"], "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "warmup_actual": main["warmup_actual"], "ms_per_step": main["ms_per_step"], "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "value_is": "camera views trained per second over all GPUs (= optimizer steps/s x n_gpus: every rank "
                        "renders one view per iteration and all ranks apply the same all-reduced update)",
            "optimizer_steps_per_s": main["optimizer_steps_per_s"], "views_per_s": main["views_per_s"],
            "config": {"workload": c["workload"], "gaussians": c["n"], "width": c["w"], "height": c["h"],
                       "views": c["views"], "views_per_iter_per_gpu": main["views_per_iter_per_gpu"],
                       "global_views_per_iter": world * main["views_per_iter_per_gpu"],
                       "execution": ("one CUDA graph replay per iteration (static-capacity intersection lists, "
                                     "no host sync); every kernel of the eager step runs in every replay"
                                     if graph_mode else "eager launches"),
                       "graph": main["graph"],
                       "fused_outputs": main["fused_outputs"],  # dn_step.py: activations / SH concat / image glue
                       # inside our kernels (True) or as the reference's torch ops (False)
                       "prune_lists": main["prune_lists"],
                       # FSB_STEP_METRICS=1: get_metrics_dict (PSNR / SSIM / depth metrics, fused, no host sync) runs
                       # inside every iteration as under nerfstudio's Trainer
                       "step_metrics": main["step_metrics"],
                       # what the e2e leg copies per view: the 8-bit RGB / normal images as a dataset holds them (the
                       # x / 255.0 of get_gt_img / dn_dataset.py:205 runs on the device), float32 depth
                       "host_targets": main["host_targets"],
                       "grad_exchange": (None if world == 1 else
                                         ("own kernels over NVLink peer memory (pack, barrier, in-place all-reduce in chunks "
                                          "on a high-priority stream, Adam of chunk k beside the all-reduce of chunk "
                                          "k + 1), captured in the step's graph"
                                          if ctx.exchange == "peer" else
                                          "one flat NCCL all-reduce captured in the step's graph")),
                       "l2": "per-step working set (parameters, Adam state, gradients, intersection lists, images: "
                             ">1 GB touched per step at cfg4, >400 MB at cfg2) exceeds the 126 MB L2; no explicit "
                             "flush"},
            "e2e": main["e2e"],
            "gpu_launches": main["gpu_launches"],
            "gpu_launches_per_step": main["gpu_launches_per_step"],
            "clocks": main["clocks"],
            "host": {"cpus_bound_to_gpu_numa_node": ctx.cpu_affinity},
            "roofline": main.get("roofline"),
        }
        if secondary is not None:
            line["secondary"] = {
                "workload": CONFIGS["cfg2"]["workload"], "value": secondary["value"], "unit": UNIT,
                "ms_per_step": secondary["ms_per_step"], "steps": max(args.steps, 200),
                "optimizer_steps_per_s": secondary["optimizer_steps_per_s"], "e2e": secondary["e2e"],
                "shim_only_eager": secondary.get("shim_only_eager"), "with_refinement": secondary.get("with_refinement"),
                "graph": secondary["graph"],
                "clocks": secondary["clocks"], "gpu_launches_per_step": secondary["gpu_launches_per_step"],
            }
    if world > 1:
        # the line first, teardown after: rank 0 prints, everybody meets, and the processes leave without
        # destroy_process_group (tearing down a communicator whose collectives sit in captured graphs hung a 2-GPU
        # tool run until its timeout, r02g)
        if line is not None:
            line["cpu_baseline"] = None
            _emit(line)
        dist.barrier()
        torch.cuda.synchronize()
        sys.stdout.flush(); sys.stderr.flush()
        os._exit(0)
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
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", default="cfg4", choices=sorted(CONFIGS))
    ap.add_argument("--mode", default="graph", choices=["graph", "eager"],
                    help="graph: the iteration is one CUDA graph replay (default); eager: per-kernel launches")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-secondary", action="store_true", help="skip the cfg2 leg")
    args = ap.parse_args()
    _quiet_stdout()
    if args.impl == "reference":
        run_reference(args)
        return
    args.warmup = max(args.warmup, 3)
    line = run_ours(args)
    if line is None:
        return
    if args.config == "cfg5":
        line["cpu_baseline"] = None  # the oracle needs ~10 min per bounded sample at 3M Gaussians: reported for cfg4 / cfg2
    elif line["n_gpus"] == 1 and not args.no_cpu_baseline:
        if _FULL_AFFINITY:
            os.sched_setaffinity(0, _FULL_AFFINITY)  # the GPU legs ran bound to the GPU's NUMA node
        sec, cores, sample = cpu_step_seconds(args.config, 1, warmup=0)
        line["cpu_baseline"] = {"value": 1.0 / sec, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample}
    else:
        line["cpu_baseline"] = None
    _emit(line)


if __name__ == "__main__":
    main()
