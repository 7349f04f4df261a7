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

All collectives go through `torch.distributed` (NCCL over NVLink on the GPU box, gloo in the CPU tests); nothing
here launches a kernel of its own.
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
