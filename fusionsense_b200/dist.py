"""Multi-GPU exchange steps of the DN-Splatter path: one process per GPU, full Gaussian replica per rank,
cameras sharded across ranks.

The reference wraps the model in DDP (/root/reference/dn_splatter/dn_pipeline.py:161-167), which cannot follow
densification (parameters are replaced every `refine_every` steps).  What keeps replicas identical here:

  * `GradSync`            one flat all-reduce (SUM) of every Gaussian parameter gradient (59 floats per Gaussian)
                          between backward and Adam; the loss is pre-scaled by 1 / world so the sum is the mean
                          over the step's global camera batch.  The captured step's overflow flag rides along.
  * `sync_densify_stats`  before `refinement_after`: SUM of `xys_grad_norm` and of the visibility increments,
                          MAX of `max_2Dsize` (12 bytes per Gaussian, every `refine_every` steps), so every rank
                          classifies split / dup / cull identically.
  * `split_generator`     the split children's normal draws come from a generator seeded by (seed, step) on every
                          rank, so the new Gaussians are bit-identical without a parameter broadcast.

  * `PeerGradExchange`    the same exchange as kernels of this library over NVLink peer memory (csrc/grad_exchange.cu):
                          pack -> barrier -> reduce-scatter -> barrier, and an Adam whose gradient load gathers the
                          reduced slices from their owners; no NCCL call on the step's path, capturable in the step's
                          CUDA graph.  torch symmetric memory provides the peer mappings (plumbing).

`GradSync`, the densification statistics and the hull slabs go through `torch.distributed` (NCCL over NVLink on the
GPU box, gloo in the CPU tests).
"""
from __future__ import annotations

from typing import Iterable, List, Optional

import torch
import torch.distributed as dist
from torch import Tensor


def world_size(group=None) -> int:
    return dist.get_world_size(group) if dist.is_available() and dist.is_initialized() else 1


def shard_views(step: int, rank: int, world: int, n_views: int, views_per_rank: int = 1) -> List[int]:
    """Camera indices rank `rank` renders at iteration `step`: the step's global batch of world * views_per_rank
    consecutive views (the reference pops them sequentially, dn_datamanager.py:100-102), dealt round-robin."""
    base = step * world * views_per_rank
    return [(base + j * world + rank) % n_views for j in range(views_per_rank)]


class GradSync:
    """Flat gradient all-reduce for a fixed list of parameters.

    `__call__(params, overflow=None)`: packs `p.grad` of every parameter (and the int32 overflow flag of the
    static-capacity step, if given) into one buffer, all-reduces it (SUM) and re-points every `p.grad` at its slice
    of the reduced buffer, so the optimizer reads it without a copy back.  Safe inside CUDA-graph capture (NCCL
    collectives are capturable; the buffer is allocated once per parameter layout)."""

    def __init__(self, group=None):
        self.group = group
        self._flat: Optional[Tensor] = None
        self._layout = None

    def _buffer(self, params: List[Tensor], extra: int) -> Tensor:
        layout = (tuple(p.numel() for p in params), extra, params[0].device, params[0].dtype)
        if self._flat is None or self._layout != layout:
            total = sum(layout[0]) + extra
            self._flat = torch.empty(total, dtype=params[0].dtype, device=params[0].device)
            self._layout = layout
        return self._flat

    def reduce(self, grads: List[Tensor], overflow: Optional[Tensor] = None) -> List[Tensor]:
        """Pack `grads` (+ the overflow flag) into the flat buffer, all-reduce it (SUM) and return views of the
        reduced buffer shaped like `grads`; `overflow` is overwritten with "any rank overflowed"."""
        flat = self._buffer(grads, 1 if overflow is not None else 0)
        parts = [g.reshape(-1) for g in grads]
        if overflow is not None:
            parts.append(overflow.reshape(-1)[:1].to(flat.dtype))
        torch.cat(parts, out=flat)
        if world_size(self.group) > 1:
            dist.all_reduce(flat, group=self.group)
        views, o = [], 0
        for g in grads:
            n = g.numel()
            views.append(flat[o:o + n].view_as(g))
            o += n
        if overflow is not None:
            overflow.copy_(flat[o:o + 1] > 0)
        return views

    def __call__(self, params: Iterable[torch.nn.Parameter], overflow: Optional[Tensor] = None) -> None:
        params = [p for p in params if p.grad is not None]
        if world_size(self.group) == 1 or not params:
            return
        for p, v in zip(params, self.reduce([p.grad for p in params], overflow)):
            p.grad = v


class PeerGradExchange:
    """Gradient exchange over NVLink peer memory with this library's own kernels (include/fsb200.h, fsb_xchg_*).

    Rank r owns slice r of the flat gradient (59 floats per Gaussian, tensor offsets rounded to 4 floats).
    `exchange(grads, overflow)` — inside or outside a CUDA-graph capture — packs this rank's gradients into its
    symmetric buffer, meets the other ranks (the overflow flag is OR-ed on the way), all-reduces IN PLACE — each rank
    reduces its own slice out of every replica and writes the sum back into every replica (NVSwitch
    multimem.ld_reduce + multimem.st where the buffers have a multicast mapping, else peer loads + peer stores) — and
    meets them again.  `adam_args()` then points CapturedAdam at this rank's own buffer.
    `mode="gather"` (FSB_XCHG_MODE=gather) is the first form: reduce-scatter into a slice buffer, and Adam reads the
    reduced slices out of their owners' memory (slower at 8 GPUs: 1.6 ms against 0.7 ms for NCCL, r02m).

    Every element is reduced once, by its owner, so all replicas apply bit-identical updates."""

    capturable = True
    MAX_CHUNKS = 8
    N_SLOTS = 2 + MAX_CHUNKS

    def __init__(self, group=None, multicast: Optional[bool] = None):
        import os

        self.group = group if group is not None else dist.group.WORLD
        self.world, self.rank = dist.get_world_size(self.group), dist.get_rank(self.group)
        from ._abi import lib

        if self.world > lib.fsb_xchg_max_world():
            raise RuntimeError(f"PeerGradExchange handles up to {lib.fsb_xchg_max_world()} ranks (one NVSwitch box)")
        if multicast is None:
            # the switch pays with more than two replicas to add (2 GPUs, gather mode: 0.77 ms with, 0.60 ms without)
            env = os.environ.get("FSB_XCHG_MULTICAST")
            multicast = (env == "1") if env is not None else self.world > 2
        self.want_multicast = bool(multicast)
        # r02n, 2 GPUs, exchange + Adam: in place in four pipelined chunks 0.59 ms, gather 0.60 ms, in place in one
        # piece 0.71 ms, NCCL 0.79 ms; 8 GPUs: gather 1.6 ms, in place through the switch 0.94 ms in one piece
        self.mode = os.environ.get("FSB_XCHG_MODE", "inplace")
        if self.mode not in ("inplace", "gather"):
            raise ValueError(f"FSB_XCHG_MODE={self.mode}: inplace or gather")
        # in place: the flat buffer is all-reduced in `chunks` pieces and the Adam launch of piece k runs on a second
        # stream beside the all-reduce of piece k + 1 (one is NVLink-bound, the other HBM-bound)
        self.chunks = max(1, min(self.MAX_CHUNKS, int(os.environ.get("FSB_XCHG_CHUNKS", "4")))) if self.mode == "inplace" else 1
        self._side = None
        self._layout = None

    def _setup(self, grads: List[Tensor]) -> None:
        import ctypes

        import torch.distributed._symmetric_memory as symm_mem

        dev = grads[0].device
        ns = [g.numel() for g in grads]
        off = [0]
        for n in ns:
            off.append(off[-1] + (n + 3) // 4 * 4)
        # per = floats of one rank's share of one chunk; S = one rank's slice of the whole buffer when it is exchanged
        # in one piece (gather mode, or exchange() in place)
        self.per = -(-off[-1] // (self.world * self.chunks * 4)) * 4
        self.S = self.per * self.chunks
        self.total = self.S * self.world
        self.ns, self.off = ns, off
        self.G = symm_mem.empty(self.total, dtype=torch.float32, device=dev)
        self.R = symm_mem.empty(self.S, dtype=torch.float32, device=dev)
        self.pad = symm_mem.empty(self.N_SLOTS * 8, dtype=torch.int32, device=dev)
        self.G.zero_(); self.R.zero_(); self.pad.zero_()
        torch.cuda.synchronize(dev)
        hG = symm_mem.rendezvous(self.G, self.group)
        hR = symm_mem.rendezvous(self.R, self.group)
        hP = symm_mem.rendezvous(self.pad, self.group)
        self._handles = (hG, hR, hP)  # keep the mappings alive
        W = self.world
        self._g_ptrs = (ctypes.c_void_p * W)(*[int(p) for p in hG.buffer_ptrs])
        self._r_ptrs = (ctypes.c_void_p * W)(*[int(p) for p in hR.buffer_ptrs])
        self._pad_ptrs = (ctypes.c_void_p * W)(*[int(p) for p in hP.buffer_ptrs])
        self._g_self = (ctypes.c_void_p * 1)(int(self.G.data_ptr()))
        self.g_mc = None
        if self.want_multicast and getattr(hG, "has_multicast_support", False) and int(hG.multicast_ptr or 0) != 0:
            self.g_mc = int(hG.multicast_ptr)
        self.epoch = torch.zeros(self.N_SLOTS, dtype=torch.int32, device=dev)
        torch.cuda.synchronize(dev)
        dist.barrier(self.group)  # every pad is zeroed before anybody signals
        self._layout = (tuple(ns), dev)

    def exchange(self, grads: List[Tensor], overflow: Optional[Tensor] = None) -> None:
        import ctypes

        from ._abi import check, lib
        from .ops import _stream

        grads = [g.contiguous() for g in grads]
        if self._layout != (tuple(g.numel() for g in grads), grads[0].device):
            if torch.cuda.is_current_stream_capturing():
                raise RuntimeError("PeerGradExchange: the gradient layout changed inside a capture; run one eager "
                                   "step first (symmetric buffers cannot be created while capturing)")
            self._setup(grads)
        n = len(grads)
        st = _stream()
        src = (ctypes.c_void_p * n)(*[g.data_ptr() for g in grads])
        ns = (ctypes.c_int64 * n)(*self.ns)
        off = (ctypes.c_int64 * (n + 1))(*self.off)
        check(lib.fsb_xchg_pack(n, ctypes.addressof(src), ctypes.addressof(ns), ctypes.addressof(off), self.G.data_ptr(),
                                self.total, st), "fsb_xchg_pack")
        self._keep = grads
        flag = None if overflow is None else overflow.data_ptr()
        check(lib.fsb_xchg_barrier(self.world, self.rank, ctypes.addressof(self._pad_ptrs), 0, self.epoch.data_ptr(),
                                   flag, st), "fsb_xchg_barrier")
        if self.mode == "inplace":
            check(lib.fsb_xchg_allreduce(self.world, self.rank, ctypes.addressof(self._g_ptrs), self.g_mc, self.S, st),
                  "fsb_xchg_allreduce")
        else:
            check(lib.fsb_xchg_reduce_scatter(self.world, self.rank, ctypes.addressof(self._g_ptrs), self.g_mc, self.S,
                                              self.R.data_ptr(), st), "fsb_xchg_reduce_scatter")
        # "my slice is reduced (and, in place: written into every replica)" — and every rank has finished reading G, so
        # the next step may overwrite it.  gather mode: the slices themselves are safe until barrier 0 of the next
        # step, which a rank enters only after its Adam has read them.
        check(lib.fsb_xchg_barrier(self.world, self.rank, ctypes.addressof(self._pad_ptrs), 1, self.epoch.data_ptr(),
                                   None, st), "fsb_xchg_barrier")

    def exchange_and_adam(self, grads: List[Tensor], overflow: Optional[Tensor], adam) -> None:
        """In-place mode, pipelined: pack on the current stream; barrier, then per chunk k { all-reduce of chunk k;
        barrier } on a high-priority communication stream; `adam.launch_range(chunk k)` on the current stream as soon as
        chunk k is complete on every rank, i.e. beside the all-reduce of chunk k + 1.  The current stream has waited for
        the whole exchange when this returns.  Same result as exchange() + adam.launch_xchg()."""
        import ctypes

        from ._abi import check, lib
        from .ops import _stream

        if self.mode != "inplace" or self.chunks == 1:
            self.exchange(grads, overflow)
            adam.launch_xchg(self, skip_flag=overflow)
            return
        grads = [g.contiguous() for g in grads]
        if self._layout != (tuple(g.numel() for g in grads), grads[0].device):
            if torch.cuda.is_current_stream_capturing():
                raise RuntimeError("PeerGradExchange: the gradient layout changed inside a capture; run one eager "
                                   "step first (symmetric buffers cannot be created while capturing)")
            self._setup(grads)
        n = len(grads)
        st = _stream()
        src = (ctypes.c_void_p * n)(*[g.data_ptr() for g in grads])
        ns = (ctypes.c_int64 * n)(*self.ns)
        off = (ctypes.c_int64 * (n + 1))(*self.off)
        check(lib.fsb_xchg_pack(n, ctypes.addressof(src), ctypes.addressof(ns), ctypes.addressof(off), self.G.data_ptr(),
                                self.total, st), "fsb_xchg_pack")
        self._keep = grads
        flag = None if overflow is None else overflow.data_ptr()
        # The communication runs on a HIGH-PRIORITY stream of its own: barrier, then per chunk { all-reduce; barrier;
        # event }, while the current stream waits for each chunk's event and launches that chunk's Adam.  Without the
        # priority the block scheduler lets whichever of Adam(k) / all-reduce(k + 1) was launched first fill the SMs, and
        # Adam's thousands of CTAs starve the next all-reduce (r02n, 2 GPUs: four chunks on equal-priority streams
        # were slower than one, 0.76 against 0.72 ms).
        main = torch.cuda.current_stream()
        if self._side is None:
            self._side = torch.cuda.Stream(device=grads[0].device, priority=-1)
        comm = self._side
        packed = torch.cuda.Event()
        packed.record(main)
        comm.wait_event(packed)
        W = self.world
        chunk = self.per * W  # floats per chunk
        events = []
        with torch.cuda.stream(comm):
            cst = _stream()
            check(lib.fsb_xchg_barrier(W, self.rank, ctypes.addressof(self._pad_ptrs), 0, self.epoch.data_ptr(), flag, cst),
                  "fsb_xchg_barrier")
            for k in range(self.chunks):
                shift = k * chunk * 4  # bytes
                ptrs = (ctypes.c_void_p * W)(*[int(p) + shift for p in self._g_ptrs])
                mc = None if self.g_mc is None else self.g_mc + shift
                check(lib.fsb_xchg_allreduce(W, self.rank, ctypes.addressof(ptrs), mc, self.per, cst), "fsb_xchg_allreduce")
                check(lib.fsb_xchg_barrier(W, self.rank, ctypes.addressof(self._pad_ptrs), 2 + k, self.epoch.data_ptr(),
                                           None, cst), "fsb_xchg_barrier")
                ev = torch.cuda.Event()
                ev.record(comm)
                events.append(ev)
        for k, ev in enumerate(events):
            main.wait_event(ev)
            adam.launch_range(self, k * chunk, (k + 1) * chunk, skip_flag=overflow)

    def probe(self, device) -> None:
        """Collective: establish a tiny symmetric buffer, so that a box without peer-memory support fails here (before
        any capture) rather than in the first step."""
        import torch.distributed._symmetric_memory as symm_mem

        t = symm_mem.empty(64, dtype=torch.float32, device=device)
        t.zero_()
        h = symm_mem.rendezvous(t, self.group)
        if len(h.buffer_ptrs) != self.world:
            raise RuntimeError("symmetric memory rendezvous returned the wrong number of peers")

    def adam_args(self):
        """(flat offsets, world, host array of the reduced slices' peer pointers, S) for fsb_adam_multi_xchg; in place:
        one "slice" — this rank's own, fully reduced buffer."""
        if self.mode == "inplace":
            return self.off, 1, self._g_self, self.total
        return self.off, self.world, self._r_ptrs, self.S

    def reduced_flat(self) -> Tensor:
        """Test helper: the whole reduced gradient gathered from the owners (a copy)."""
        if self.mode == "inplace":
            return self.G.clone()
        hR = self._handles[1]
        parts = [hR.get_buffer(w, (self.S,), torch.float32).clone() for w in range(self.world)]
        return torch.cat(parts)


@torch.no_grad()
def sync_densify_stats(model, group=None) -> None:
    """All ranks end up with the statistics a single process would have gathered over the global camera batch.

    splatfacto's after_train starts `vis_counts` at ones and adds one per visible step, so the
    increments (vis_counts - 1) are summed, not the counters.  A rank that has not accumulated anything since the
    last refinement contributes zeros."""
    if world_size(group) == 1:
        return
    n = model.num_points
    dev = model.gauss_params["means"].device
    zeros = lambda: torch.zeros(n, dtype=torch.float32, device=dev)  # noqa: E731
    grad_norm = model.xys_grad_norm if model.xys_grad_norm is not None else zeros()
    vis_inc = (model.vis_counts - 1.0) if model.vis_counts is not None else zeros()
    max2d = model.max_2Dsize if model.max_2Dsize is not None else zeros()
    sums = torch.stack([grad_norm, vis_inc])
    dist.all_reduce(sums, group=group)
    max2d = max2d.clone()
    dist.all_reduce(max2d, op=dist.ReduceOp.MAX, group=group)
    model.xys_grad_norm = sums[0].contiguous()
    model.vis_counts = (sums[1] + 1.0).contiguous()
    model.max_2Dsize = max2d


def split_generator(seed: int, step: int, device) -> torch.Generator:
    """The generator every rank draws the split children from at refinement step `step`."""
    g = torch.Generator(device=device)
    g.manual_seed((int(seed) * 1_000_003 + int(step)) & 0x7FFFFFFFFFFFFFFF)
    return g


@torch.no_grad()
def synchronised_refinement(model, optimizers, step: int, seed: int = 0, group=None):
    """`refinement_after` on every rank with identical inputs -> identical replicas (no parameter broadcast)."""
    from .densify import refinement_after

    sync_densify_stats(model, group)
    gen = split_generator(seed, step, model.gauss_params["means"].device)
    return refinement_after(model, optimizers, step, generator=gen)


@torch.no_grad()
def replicas_identical(params: Iterable[Tensor], group=None) -> bool:
    """Debug / test helper: every rank holds bit-identical parameters (compares against rank 0's copy)."""
    ok = True
    for p in params:
        ref = p.detach().clone()
        dist.broadcast(ref, src=0, group=group)
        ok = ok and bool(torch.equal(ref, p.detach()))
    flag = torch.tensor([1 if ok else 0], device=ref.device, dtype=torch.int32)
    dist.all_reduce(flag, op=dist.ReduceOp.MIN, group=group)
    return bool(flag.item())
