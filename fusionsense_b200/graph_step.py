"""The DN-Splatter training iteration as ONE CUDA graph launch.

At FusionSense's scene sizes (300 k Gaussians, 640x480) the eager iteration is host-bound: ~135 launches per step
(40 of ours, the rest the torch glue `dn_model.py` issues around the gsplat calls) and one stream synchronisation
for the intersection count keep the GPU ~45 % busy (profiles/).  The kernels' static-capacity mode
(include/fsb200.h) removes the synchronisation and fixes the launch sequence, so the whole iteration

    zero_grad -> get_outputs (rasterization RGB+ED, normals, legacy normals pass) -> get_loss_dict -> backward
    -> Adam (all groups, one launch) -> after_train statistics

is captured once with `torch.cuda.graph` and replayed per step; the per-replay inputs (camera index, Adam step
sizes) are uploaded with fsb_upload_small.  Nothing is skipped or cached between replays: every kernel of the
eager step runs in every replay on the live parameters.

Re-capture happens when something the capture froze changes: the number of Gaussians (refinement every
`refine_every` steps), the active SH degree, the binary-opacity window, the end of densification, or an overflow of
the intersection capacity (the device then leaves parameters, Adam state and statistics untouched for that step;
the host notices through `poll()`, grows the capacity and re-runs).
"""
from __future__ import annotations

import ctypes
import os
import time
from typing import Dict, Optional, Tuple

import torch
from torch import Tensor

from . import ops
from ._abi import check, lib
from .dn_step import DNSplatterStep
from .optim import CapturedAdam


U8_TARGET_KEYS = ("image", "normal")


def eight_bit_targets(batch: Dict[str, Tensor]) -> Tuple[Dict[str, Tensor], Dict[str, Tensor]]:
    """A view's targets as a FusionSense dataset holds them: RGB and normal map as 8-bit images (images/rgb_i.png,
    normals_from_pretrain/*.png), depth as float32.  Returns (resident, host): `host` keeps the two images as pinned
    uint8 (what nerfstudio's image cache holds and what crosses PCIe), `resident` is what the reference's loaders make
    of those bytes on the device — `image.float() / 255.0` (splatfacto get_gt_img) and numpy's
    `normal_map.astype("float32") / 255.0` (dn_dataset.py:205) — computed by the same kernel `stage_async` runs after
    the copy, so a staged view and a resident one are bit-identical."""
    from .compose import u8_to_unit_float

    resident, host = {}, {}
    for k, t in batch.items():
        if k in U8_TARGET_KEYS:
            q = torch.round(t.detach().clamp(0, 1) * 255.0).to(torch.uint8).contiguous()
            resident[k] = u8_to_unit_float(q.view(-1), recip=(k != "normal")).view(t.shape)
            host[k] = q.cpu().pin_memory()
        else:
            resident[k] = t
            host[k] = t.detach().cpu().pin_memory()
    return resident, host


class GraphedDNSplatterStep:
    def __init__(self, model: DNSplatterStep, targets: Dict[int, Dict[str, Tensor]], capacity: Optional[int] = None,
                 margin: float = 1.3, grad_sync=None, loss_scale: float = 1.0, views_per_iter: int = 1):
        """`targets[v]` = batch of view v (`image`, `sensor_depth`, `normal`, device tensors).  They are stacked
        into one resident tensor per key; `stage(v, host_batch)` overwrites a view's slot from host memory.
        `grad_sync(params, overflow_flag)`: optional hook run inside the captured step between backward and Adam
        (multi-GPU: the NCCL all-reduce of the parameter gradients and of the overflow flag); `loss_scale`
        multiplies the loss before backward (1 / world_size keeps the mean over the global camera batch).
        `views_per_iter` = V > 1: a camera batch per iteration (BASELINE.json configs[4]: 32 views per step, dealt over
        the ranks): the V views are rendered and differentiated one after the other inside the one captured step,
        their gradients accumulate (each loss scaled by 1 / V), the densification statistics take every view, and
        Adam steps once.  `train_iteration` then takes V view indices."""
        if model.device.type != "cuda":
            raise RuntimeError("GraphedDNSplatterStep needs a CUDA model (no CPU fallback)")
        if not model._fused_optim:
            raise RuntimeError("GraphedDNSplatterStep needs fused_optimizer=True")
        self.model = model
        self.device = model.device
        views = sorted(targets)
        assert views == list(range(len(views))), "targets must cover views 0..V-1"
        self.targets = {k: torch.stack([targets[v][k] for v in views]).contiguous() for k in targets[views[0]]}
        self.margin = float(margin)
        self.capacity = int(capacity) if capacity else None
        self.views_per_iter = V = max(1, int(views_per_iter))
        self.cam = torch.zeros(V, dtype=torch.int64, device=self.device)
        self._cam_host = (ctypes.c_int64 * V)()
        self.overflow = torch.zeros(1, dtype=torch.int32, device=self.device)
        # [loss, overflowed steps so far, n_isects of the RGB+ED pass, n_isects of the legacy normals pass]
        self.result = torch.zeros(4, dtype=torch.float64, device=self.device)
        self._overflow_seen = 0
        self.max_isects_seen = 0
        self.graph: Optional[torch.cuda.CUDAGraph] = None
        self.graph_tail: Optional[torch.cuda.CUDAGraph] = None  # N > 1: Adam + statistics, after the eager all-reduce
        self.signature = None
        self.adam: Optional[CapturedAdam] = None
        self.captures = 0
        self.capture_seconds = 0.0  # host wall time spent inside capture() (warm-up iteration + capture), cumulative
        self.replays = 0
        self.grad_sync = grad_sync
        # N > 1: capture the gradient exchange inside the one graph (True) or launch it eagerly between two graphs
        import os as _os

        # (captured by default: r02e ran it at N = 2 without the round-1 hang, 1.15 ms per cfg2 step against 1.36 ms with
        # two graphs; FSB_CAPTURE_NCCL=0 brings the two-graph form back)
        self.capture_collective = _os.environ.get("FSB_CAPTURE_NCCL", "1") != "0"
        # a dist.PeerGradExchange: the exchange is this library's kernels over NVLink peer memory, folded into Adam
        self._peer = grad_sync is not None and hasattr(grad_sync, "exchange")
        self.loss_scale = float(loss_scale)
        self.launches_per_replay = 0  # libfsb200 kernels inside one replay (counted while capturing)
        # overflowed steps are no-ops on the device but advance the host's step counters; callers that never poll()
        # would keep replaying with a capacity that is too small, so every `check_every` replays the result vector
        # is copied to pinned memory and looked at once the copy has landed (no stream synchronisation)
        self.check_every = 16
        self._last_check = 0
        self._check_slot = None
        self._copy_stream = None
        self._slot_staged: Dict[int, torch.cuda.Event] = {}  # view -> "prefetch finished"
        self._slot_read: Dict[int, torch.cuda.Event] = {}    # view -> "last replay that read the slot finished"
        self._host_ring = None
        self._ev_prev = self._ev_last = None  # end of the replay before the most recent one / of the most recent one

    # ---- what the capture freezes ------------------------------------------------------------
    def _signature(self):
        m, cfg = self.model, self.model.config
        skip_steps = cfg.reset_alpha_every * cfg.refine_every
        binary = (cfg.use_binary_opacities and m.step > cfg.warmup_length and not m.step % skip_steps == 0
                  and m.step % skip_steps not in range(1, 200 + 1))
        # the densification statistics are written by the captured fsb_densify_stats: refinement_after drops them
        # (None) on EVERY refine step, also on the ones that neither densify nor cull, so their identity is part of
        # what the capture froze (round-1 advisor finding: the graph kept writing into the freed tensors)
        stats = tuple(id(t) if t is not None else None for t in (m.xys_grad_norm, m.vis_counts, m.max_2Dsize))
        return (m.num_points, min(m.step // cfg.sh_degree_interval, cfg.sh_degree), binary,
                m.step >= cfg.stop_split_at, self.capacity, tuple(id(p) for p in m.gauss_params.values()), stats)

    @torch.no_grad()
    def _probe_capacity(self) -> int:
        """Largest intersection count over the training views (legacy bbox rule, the larger of the two), by the
        eager path with its one host read per view."""
        m = self.model
        worst = 0
        training, m.training = m.training, False
        try:
            for v in range(self.targets["image"].shape[0]):
                m.get_outputs(v)
                from .gsplat.cuda_legacy import _wrapper as legacy

                worst = max(worst, int(legacy._LAST_BINNING.get("n_isects", 0)))
        finally:
            m.training = training
        return worst

    def _body_main(self, warmup: bool = False):
        """zero_grad -> get_outputs -> get_loss_dict -> backward in static-capacity mode on the current stream.
        `warmup`: run eagerly before the capture with the skip flag raised, so every lazy initialisation happens
        outside the graph while parameters, Adam state and statistics stay untouched."""
        m = self.model
        for opt in m.optimizers.values():
            opt.zero_grad(set_to_none=True)
        if warmup:
            self.overflow.fill_(1)
        else:
            self.overflow.zero_()
        with ops.static_capacity(self.capacity, self.overflow) as st:
            V = self.views_per_iter
            total = None
            for j in range(V):
                cam_j = self.cam[j:j + 1]
                # this view's targets are gathered on a side stream while the main one projects, bins and composites:
                # three image-sized copies (8.6 MB at 640x480) that nothing needs before the losses
                cur = torch.cuda.current_stream()
                if getattr(self, "_select_stream", None) is None:
                    self._select_stream = torch.cuda.Stream(device=self.device)
                sel = self._select_stream
                sel.wait_stream(cur)
                with torch.cuda.stream(sel):
                    batch = {k: t.index_select(0, cam_j)[0] for k, t in self.targets.items()}
                outputs = m.get_outputs(cam_j)
                cur.wait_stream(sel)
                for t in batch.values():
                    t.record_stream(cur)
                if m.config.step_metrics:
                    m.last_metrics = m.get_metrics_dict(outputs, batch)  # device tensors, rewritten by every replay
                loss_dict = m.get_loss_dict(outputs, batch)
                loss = loss_dict["main_loss"] + loss_dict["scale_reg"]
                scale = self.loss_scale / V
                (loss * scale if scale != 1.0 else loss).backward()
                total = loss.detach() if total is None else total + loss.detach()
                if j < V - 1:
                    # statistics of every view but the last here (they need this view's radii / absgrad); the last
                    # view's follow the gradient exchange in _body_tail, where the overflow flag covers all ranks
                    m.after_train(skip_flag=self.overflow)
            loss = total / V if V > 1 else total
            counts = list(st.counts)
        return loss, counts

    def _body_tail(self, loss, counts, warmup: bool = False):
        """Adam (all groups, one launch) -> after_train statistics -> result vector."""
        m = self.model
        if self._peer:
            # gradients of the Adam entries, in its order, through our own NVLink exchange; Adam gathers the result
            grads = [p.grad for _, _, p in self.adam.entries]
            self.grad_sync.exchange_and_adam(grads, self.overflow, self.adam)
        else:
            self.adam.launch(skip_flag=self.overflow)
        m.after_train(skip_flag=self.overflow)
        if warmup:
            return
        with torch.no_grad():
            self.result[0:1].copy_(loss.reshape(1))
            self.result[1:2].add_(self.overflow)
            for i, c in enumerate(counts[:2]):
                self.result[2 + i:3 + i].copy_(c)

    def _sync_grads(self):
        """Multi-GPU exchange step between backward and Adam, launched eagerly on the current stream: the
        gradients the captured backward left in its (address-stable) buffers are packed with the overflow flag,
        all-reduced, and handed to Adam as views of the reduced buffer."""
        views = self.grad_sync.reduce(self._grad_src, self.overflow)
        return views

    def capture(self) -> None:
        """The whole iteration is one graph; with a `grad_sync` (N > 1) the gradient exchange is captured between
        backward and Adam.  (`capture_collective=False`: two graphs with the exchange launched eagerly between
        them, the round-1 form.)"""
        m = self.model
        if self.capacity is None or self.capacity <= 0:
            seen = self.max_isects_seen or self._probe_capacity()
            self.capacity = int(seen * self.margin) + 65536
        if m.xys_grad_norm is None:
            m.xys_grad_norm = torch.zeros(m.num_points, device=self.device, dtype=torch.float32)
            m.vis_counts = torch.ones(m.num_points, device=self.device, dtype=torch.float32)
        if m.max_2Dsize is None:
            m.max_2Dsize = torch.zeros(m.num_points, device=self.device, dtype=torch.float32)
        self.adam = CapturedAdam(m.optimizers.values())
        # (A private pool shared across re-captures does not work: dropping the only graph of a pool releases the pool,
        # and capturing into the stale handle trips an allocator assert — r02s.  Each capture takes a fresh pool; the
        # cost of a re-capture, ~0.18 s at cfg2, is reported by bench.py as `with_refinement.host_seconds`.)
        self.graph = self.graph_tail = None
        self._keep = None
        params = [p for p in m.gauss_params.values()]
        side = self._side = getattr(self, "_side", None) or torch.cuda.Stream()  # warm-up and capture share it
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            loss, counts = self._body_main(warmup=True)
            if self.grad_sync is not None and not self._peer:
                self._grad_src = [p.grad for p in params if p.grad is not None]
                for p, v in zip([p for p in params if p.grad is not None], self._sync_grads()):
                    p.grad = v
            self._body_tail(loss, counts, warmup=True)
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        # parameters that already took part in an eager backward own AccumulateGrad nodes tagged with that stream;
        # with zero_grad(set_to_none=True) those nodes only store the incoming gradient (no launch), so the
        # stream-mismatch notice torch prints for them does not apply to this capture
        torch.autograd.graph.set_warn_on_accumulate_grad_stream_mismatch(False)
        g = torch.cuda.CUDAGraph()
        n0 = lib.fsb_launch_count()
        if self._peer:
            with torch.cuda.graph(g, stream=side):
                loss, counts = self._body_main()
                self._body_tail(loss, counts)
            tail = None
        elif self.grad_sync is not None and self.capture_collective:
            # ONE graph: the gradient exchange is captured between backward and Adam
            with torch.cuda.graph(g, stream=side):
                loss, counts = self._body_main()
                live = [p for p in params if p.grad is not None]
                self._grad_src = [p.grad for p in live]
                for p, v in zip(live, self._sync_grads()):
                    p.grad = v
                self._body_tail(loss, counts)
            tail = None
        elif self.grad_sync is None:
            with torch.cuda.graph(g, stream=side):
                loss, counts = self._body_main()
                self._body_tail(loss, counts)
            tail = None
        else:
            with torch.cuda.graph(g, stream=side):
                loss, counts = self._body_main()
            live = [p for p in params if p.grad is not None]
            self._grad_src = [p.grad for p in live]  # address-stable: rewritten by every replay of `g`
            with torch.cuda.stream(side):
                views = self._sync_grads()  # on capture-time garbage; fixes the reduced buffer Adam will read
            for p, v in zip(live, views):
                p.grad = v
            torch.cuda.synchronize()
            tail = torch.cuda.CUDAGraph()
            with torch.cuda.graph(tail, stream=side, pool=g.pool()):
                self._body_tail(loss, counts)
            self._keep = (loss, counts)  # read by the tail graph: keep the main graph's buffers alive
        self.launches_per_replay = int(lib.fsb_launch_count() - n0)
        self.graph = g
        self.graph_tail = tail
        self.signature = self._signature()
        self.captures += 1

    # ---- per step ----------------------------------------------------------------------------
    def train_iteration(self, cam_idx) -> Tensor:
        """One training iteration on view `cam_idx` (a sequence of `views_per_iter` views when that is > 1); returns
        the device-resident result vector [loss, overflowed steps, n_isects, n_isects (normals pass)] (float64,
        overwritten by the next call)."""
        m = self.model
        cams = [int(cam_idx)] if self.views_per_iter == 1 and not isinstance(cam_idx, (list, tuple)) else [int(c) for c in cam_idx]
        if len(cams) != self.views_per_iter:
            raise ValueError(f"train_iteration takes {self.views_per_iter} view indices, got {len(cams)}")
        if self.graph is None or self.signature != self._signature():
            t0 = time.perf_counter()
            self.capture()
            self.capture_seconds += time.perf_counter() - t0
        m.optimizers["means"].param_groups[0]["lr"] = m._means_lr()
        self.adam.advance()
        for j, c in enumerate(cams):
            self._cam_host[j] = c
        check(lib.fsb_upload_small(self.cam.data_ptr(), ctypes.addressof(self._cam_host), 8 * len(cams), ops._stream()),
              "fsb_upload_small")
        for c in cams:
            staged = self._slot_staged.pop(c, None)
            if staged is not None:
                torch.cuda.current_stream().wait_event(staged)
        self.graph.replay()
        self._ev_prev, self._ev_last = self._ev_last, torch.cuda.Event()
        self._ev_last.record()
        for c in cams:
            self._slot_read[c] = self._ev_last
        if self.graph_tail is not None:
            self._sync_grads()
            self.graph_tail.replay()
        m.step += 1
        self.replays += 1
        if self.replays - self._last_check >= self.check_every:
            self._check_async()
        return self.result

    # ---- pipelined host <-> device traffic (the e2e leg of bench.py) ---------------------------------------
    def stage_async(self, cam_idx: int, host_batch: Dict[str, Tensor]) -> int:
        """Prefetch a view's targets from pinned host memory into its resident slot on a copy stream, so the copy
        of the NEXT step's inputs overlaps this step's kernels.  `train_iteration(cam_idx)` waits for it.  The slot
        must not be the one a replay in flight reads (callers prefetch the following step's view); the copy is
        ordered after the last replay that read the slot."""
        if self._copy_stream is None:
            # high priority: the byte -> float conversions queued behind the copies are a few microseconds of work and
            # must not wait for the step's 28k-CTA kernels to drain before the next copy can start
            self._copy_stream = torch.cuda.Stream(device=self.device, priority=-1)
        cs = self._copy_stream
        # the slot was last read by `_slot_read[cam_idx]`; a slot no tracked replay has read can only have been
        # touched before the replay launched most recently (that one reads another view's slot)
        ev = self._slot_read.get(cam_idx) or self._ev_prev
        if ev is not None:
            cs.wait_event(ev)
        else:
            cs.wait_stream(torch.cuda.current_stream())
        n = 0
        with torch.cuda.stream(cs):
            pending = []
            for k, t in host_batch.items():
                n += self._copy_target(k, cam_idx, t, convert_later=pending)
            for job in pending:  # all copies are queued before the first conversion kernel: the copy engine never idles
                job()
            done = torch.cuda.Event()
            done.record(cs)
        self._slot_staged[cam_idx] = done
        return n

    def read_result_async(self):
        """Queue a 32-byte device -> host copy of this step's result vector into a pinned ring slot and return the
        PREVIOUS step's values (or None on the first call): the host reads every step's loss while staying one
        step ahead of the device instead of draining the stream each iteration."""
        if self._host_ring is None:
            self._host_ring = [torch.zeros(4, dtype=torch.float64).pin_memory() for _ in range(2)]
            self._host_events = [None, None]
            self._ring_i = 0
        i = self._ring_i
        self._host_ring[i].copy_(self.result, non_blocking=True)
        ev = torch.cuda.Event()
        ev.record()
        self._host_events[i] = ev
        prev = None
        j = 1 - i
        if self._host_events[j] is not None:
            self._host_events[j].synchronize()
            prev = self._host_ring[j].tolist()
            self.max_isects_seen = max(self.max_isects_seen, int(prev[2]), int(prev[3]))
        self._ring_i = j
        return prev

    def stage(self, cam_idx: int, host_batch: Dict[str, Tensor]) -> int:
        """Copy a view's targets from (pinned) host memory into its resident slot; returns the bytes moved."""
        n = 0
        for k, t in host_batch.items():
            n += self._copy_target(k, cam_idx, t)
        return n

    def _copy_target(self, key: str, cam_idx: int, t: Tensor, convert_later: Optional[list] = None) -> int:
        """Host tensor -> resident slot on the current stream; returns the bytes that crossed PCIe.  uint8 sources
        (the 8-bit RGB / normal images as they are on disk) travel as bytes into a device staging buffer and are
        turned into the slot's float32 by `fsb_u8_to_unit_float` on the same stream — what the reference does with
        `image.float() / 255.0` after its own H2D copy (splatfacto get_gt_img; dn_dataset.py:205 for normals)."""
        slot = self.targets[key][cam_idx]
        if t.dtype == torch.uint8 and slot.dtype == torch.float32:
            from .compose import u8_to_unit_float

            if getattr(self, "_u8_stage", None) is None:
                self._u8_stage: Dict[tuple, Tensor] = {}
            sk = (key, ops._stream())  # raw handle of the current stream (no Stream object per call)
            buf = self._u8_stage.get(sk)
            if buf is None or buf.numel() != t.numel():
                # one buffer per (key, stream): copies and conversions of successive views are ordered by the stream
                buf = self._u8_stage[sk] = torch.empty(t.numel(), dtype=torch.uint8, device=self.device)
            buf.copy_(t.reshape(-1), non_blocking=True)
            recip = key != "normal"  # normals: numpy's division (dn_dataset.py:205)
            if convert_later is None:
                u8_to_unit_float(buf, slot, recip=recip)
            else:
                convert_later.append(lambda b=buf, s_=slot, r=recip: u8_to_unit_float(b, s_, recip=r))
        else:
            slot.copy_(t, non_blocking=True)
        return t.numel() * t.element_size()

    def _check_async(self) -> None:
        self._last_check = self.replays
        if self._check_slot is not None:
            host, ev = self._check_slot
            if not ev.query():
                return  # the previous probe has not landed yet: look again next time
            vals = host.tolist()
            self._check_slot = None
            self._handle(vals)
        host = torch.zeros(4, dtype=torch.float64).pin_memory() if getattr(self, "_check_host", None) is None \
            else self._check_host
        self._check_host = host
        host.copy_(self.result, non_blocking=True)
        ev = torch.cuda.Event()
        ev.record()
        self._check_slot = (host, ev)

    def _handle(self, vals):
        loss, overflowed, n0, n1 = vals
        self.max_isects_seen = max(self.max_isects_seen, int(n0), int(n1))
        new = int(overflowed) - self._overflow_seen
        if new > 0:
            self._overflow_seen = int(overflowed)
            self.model.step -= new
            self.adam.rollback(new)
            self.capacity = int(max(self.capacity * 1.5, self.max_isects_seen * self.margin)) + 65536
            self.graph = None
        return new

    def poll(self) -> Dict[str, float]:
        """Read the result vector (one 32-byte D2H copy; synchronises with the last replay) and handle overflowed
        steps: they changed nothing on the device, so the host counters are rolled back, the capacity grows and the
        next call re-captures."""
        loss, overflowed, n0, n1 = vals = self.result.tolist()
        self._check_slot = None
        new = self._handle(vals)
        return {"loss": loss, "overflowed_steps": int(overflowed), "n_isects": int(n0), "n_isects_normals": int(n1),
                "capacity": self.capacity, "new_overflows": new}
