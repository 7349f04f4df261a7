"""Densify / prune bookkeeping of the DN-Splatter model on libfsb200's kernels (csrc/refine.cu, csrc/hull_prune.cu).

Host-side mirror of
  * `DNSplatterModel.refinement_after`   /root/reference/dn_splatter/dn_model.py:326-451
  * `DNSplatterModel.hull_pruning`       /root/reference/dn_splatter/dn_model.py:1249-1276
  * `remove_from_all_optim` / `dup_in_all_optim`   dn_model.py:149-170
and of the nerfstudio 1.1.3 helpers they drive (split_gaussians, dup_gaussians, cull_gaussians, dup_in_optim,
remove_from_optim).  Same names, same step gating, same effects on `gauss_params`, on every
optimizer's `state[param]["exp_avg" / "exp_avg_sq"]` and `param_groups[0]["params"]`, and on `add_mask`.

`model` is duck-typed: anything with `gauss_params` (dict name -> nn.Parameter), `config`, `step`,
`num_train_data`, `last_size`, `xys_grad_norm`, `vis_counts`, `max_2Dsize`, `add_mask`, `device`
(DNSplatterModel itself, or fusionsense_b200.dn_step.DNSplatterStep).  `optimizers` is a dict name -> optimizer with
one parameter per optimizer (dn_config.py:36-75), i.e. nerfstudio's `Optimizers.optimizers`.

There is no CPU path: CUDA tensors only.
"""
from __future__ import annotations

import math
from typing import Dict, Optional

import torch
from torch import Tensor

from . import ops
from ._abi import check, lib, ptr
from .ops import _f32c, _req_cuda, _stream

SIZE_FAC = 1.6  # nerfstudio split_gaussians


def _u8(mask: Optional[Tensor]) -> Optional[Tensor]:
    if mask is None:
        return None
    return mask.reshape(-1).to(torch.uint8).contiguous()


def _scan(flags: Tensor):
    """exclusive int64 offsets of int32 flags, and the total (one D2H read)"""
    return ops.isect_scan(flags)


def _rebuild(model, optimizers: Dict[str, torch.optim.Optimizer], M: int, N: int, parent: Tensor, keep: Tensor,
             offsets: Tensor, n_kept: int, fixup=None) -> None:
    """Order-preserving compaction of every parameter tensor and both Adam moments of each; re-keys the optimizer
    state by the new Parameter objects like remove_from_optim / dup_in_optim do."""
    dev = parent.device
    st = _stream()
    new_params = {}
    for name, param in model.gauss_params.items():
        src = _f32c(param.detach())
        width = int(math.prod(src.shape[1:]))
        out = torch.empty((n_kept,) + tuple(src.shape[1:]), dtype=torch.float32, device=dev)
        check(lib.fsb_refine_gather(M, N, width, ptr(src), ptr(parent), ptr(keep), ptr(offsets), 0, ptr(out), st),
              "fsb_refine_gather")
        new_params[name] = out
    if fixup is not None:
        fixup(new_params)
    for name, out in new_params.items():
        old = model.gauss_params[name]
        new = torch.nn.Parameter(out, requires_grad=old.requires_grad)
        opt = optimizers.get(name) if optimizers is not None else None
        if opt is not None:
            state = opt.state.pop(old, None)
            if state is not None and "exp_avg" in state:
                for key in ("exp_avg", "exp_avg_sq"):
                    src = _f32c(state[key])
                    width = int(math.prod(src.shape[1:]))
                    mom = torch.empty((n_kept,) + tuple(src.shape[1:]), dtype=torch.float32, device=dev)
                    check(lib.fsb_refine_gather(M, N, width, ptr(src), ptr(parent), ptr(keep), ptr(offsets), 1,
                                                ptr(mom), st), "fsb_refine_gather")
                    state[key] = mom
                opt.state[new] = state
            elif state is not None:
                opt.state[new] = state
            opt.param_groups[0]["params"] = [new]
        model.gauss_params[name] = new


def _cull_only(model, optimizers, extra_cull: Optional[Tensor]) -> Tensor:
    """cull_gaussians(extra) + remove_from_all_optim: returns the deleted mask [N] bool."""
    cfg = model.config
    gp = model.gauss_params
    N = gp["means"].shape[0]
    dev = gp["means"].device
    parent = torch.empty((N,), dtype=torch.int32, device=dev)
    keep = torch.empty((N,), dtype=torch.int32, device=dev)
    toobig_on = model.step > cfg.refine_every * cfg.reset_alpha_every
    screen_on = toobig_on and model.step < cfg.stop_screen_size_at and model.max_2Dsize is not None
    check(lib.fsb_refine_keep(N, N, 0, 0, 0, None, None, None, ptr(_f32c(gp["opacities"].detach())),
                              ptr(_f32c(gp["scales"].detach())),
                              ptr(_f32c(model.max_2Dsize)) if screen_on else None, float(cfg.cull_alpha_thresh),
                              float(cfg.cull_scale_thresh) if toobig_on else 0.0,
                              float(cfg.cull_screen_size) if screen_on else 0.0, SIZE_FAC, ptr(_u8(extra_cull)),
                              ptr(parent), ptr(keep), _stream()), "fsb_refine_keep")
    offsets, n_kept = _scan(keep)
    _rebuild(model, optimizers, N, N, parent, keep, offsets, n_kept)
    deleted = keep == 0
    if getattr(model, "add_mask", None) is not None:
        model.add_mask = model.add_mask[~deleted]
    return deleted


@torch.no_grad()
def refinement_after(model, optimizers: Dict[str, torch.optim.Optimizer], step: int,
                     samples: Optional[Tensor] = None, generator: Optional[torch.Generator] = None):
    """dn_model.py:326-451.  `samples`: optional [n_split_samples * n_split, 3] standard-normal draws for the split
    children (default: torch.randn on the model's device, like the reference).  Returns the deleted mask or None."""
    assert step == model.step
    cfg = model.config
    if model.step <= cfg.warmup_length:
        return None
    gp = model.gauss_params
    _req_cuda(gp["means"])
    dev = gp["means"].device
    N = gp["means"].shape[0]
    reset_interval = cfg.reset_alpha_every * cfg.refine_every
    do_densification = (model.step < cfg.stop_split_at
                        and model.step % reset_interval > model.num_train_data + cfg.refine_every)
    deleted_mask = None
    if do_densification:
        assert model.xys_grad_norm is not None and model.vis_counts is not None and model.max_2Dsize is not None
        scales = _f32c(gp["scales"].detach())
        action = torch.empty((N,), dtype=torch.uint8, device=dev)
        split_flag = torch.empty((N,), dtype=torch.int32, device=dev)
        dup_flag = torch.empty((N,), dtype=torch.int32, device=dev)
        screen = model.step < cfg.stop_screen_size_at
        add_mask = _u8(getattr(model, "add_mask", None))
        max2d = _f32c(model.max_2Dsize)
        check(lib.fsb_refine_classify(N, ptr(_f32c(model.xys_grad_norm)), ptr(_f32c(model.vis_counts)), ptr(max2d),
                                      ptr(scales), float(max(model.last_size[0], model.last_size[1])),
                                      float(cfg.densify_grad_thresh), float(cfg.densify_size_thresh),
                                      float(cfg.split_screen_size) if screen else 0.0, SIZE_FAC, ptr(add_mask),
                                      ptr(action), ptr(split_flag), ptr(dup_flag), _stream()), "fsb_refine_classify")
        split_rank, n_split = _scan(split_flag)
        dup_rank, n_dup = _scan(dup_flag)
        split_idcs = torch.empty((max(n_split, 1),), dtype=torch.int32, device=dev)
        dup_idcs = torch.empty((max(n_dup, 1),), dtype=torch.int32, device=dev)
        check(lib.fsb_refine_index(N, ptr(action), ptr(split_rank), ptr(dup_rank), ptr(split_idcs), ptr(dup_idcs),
                                   _stream()), "fsb_refine_index")
        samps = int(cfg.n_split_samples)
        n_children = samps * n_split
        if samples is None:
            samples = torch.randn((n_children, 3), device=dev, generator=generator)
        samples = _f32c(samples.reshape(n_children, 3))
        M = N + n_children + n_dup
        parent = torch.empty((M,), dtype=torch.int32, device=dev)
        keep = torch.empty((M,), dtype=torch.int32, device=dev)
        toobig_on = model.step > cfg.refine_every * cfg.reset_alpha_every
        screen_on = toobig_on and screen
        opac = _f32c(gp["opacities"].detach())
        check(lib.fsb_refine_keep(M, N, n_split, n_dup, samps, ptr(action), ptr(split_idcs), ptr(dup_idcs), ptr(opac),
                                  ptr(scales), ptr(max2d) if screen_on else None, float(cfg.cull_alpha_thresh),
                                  float(cfg.cull_scale_thresh) if toobig_on else 0.0,
                                  float(cfg.cull_screen_size) if screen_on else 0.0, SIZE_FAC, None, ptr(parent),
                                  ptr(keep), _stream()), "fsb_refine_keep")
        offsets, n_kept = _scan(keep)
        means, quats = _f32c(gp["means"].detach()), _f32c(gp["quats"].detach())

        def fixup(new_params):
            check(lib.fsb_refine_split_fixup(n_children + n_dup, n_children, N, ptr(action), ptr(means), ptr(scales),
                                             ptr(quats), ptr(samples), ptr(parent), ptr(keep), ptr(offsets), SIZE_FAC,
                                             ptr(new_params["means"]), ptr(new_params["scales"]), _stream()),
                  "fsb_refine_split_fixup")

        _rebuild(model, optimizers, M, N, parent, keep, offsets, n_kept, fixup)
        deleted_mask = keep == 0
        if getattr(model, "add_mask", None) is not None:
            am = torch.cat([model.add_mask, torch.zeros(n_children + n_dup, dtype=model.add_mask.dtype, device=dev)])
            model.add_mask = am[~deleted_mask]
    elif model.step >= cfg.stop_split_at and cfg.continue_cull_post_densification:
        deleted_mask = _cull_only(model, optimizers, None)

    if model.step < cfg.stop_split_at and model.step % reset_interval == cfg.refine_every:
        # opacity reset (dn_model.py:428-445)
        reset_value = cfg.cull_alpha_thresh * 2.0
        op = model.gauss_params["opacities"]
        op.data = torch.clamp(op.data, max=torch.logit(torch.tensor(reset_value)).item())
        optim = optimizers["opacities"]
        param = optim.param_groups[0]["params"][0]
        param_state = optim.state[param]
        if "exp_avg" in param_state:
            param_state["exp_avg"] = torch.zeros_like(param_state["exp_avg"])
            param_state["exp_avg_sq"] = torch.zeros_like(param_state["exp_avg_sq"])
    model.xys_grad_norm = None
    model.vis_counts = None
    model.max_2Dsize = None
    return deleted_mask


@torch.no_grad()
def hull_prune_mask(means: Tensor, visual_hull: Tensor, scale_factor: float,
                    add_mask: Optional[Tensor] = None) -> Tensor:
    """dn_model.py:1254-1269 without the N_close x V distance matrix -> bool mask [N] of Gaussians to delete."""
    _req_cuda(means, visual_hull)
    means_c, hull_c = _f32c(means.detach()), _f32c(visual_hull)
    N, V = means_c.shape[0], hull_c.shape[0]
    dev = means.device
    center = hull_c.mean(dim=0).contiguous()
    lo, hi = 0.005 * scale_factor, 0.02 * scale_factor
    min_dist = torch.empty((N,), dtype=torch.float32, device=dev)
    # no early stop (stop_below = 0): the exact minimum also serves callers that want the distances
    check(lib.fsb_hull_min_dist(N, ptr(means_c), V, ptr(hull_c), ptr(center), 0.2 * scale_factor, 0.0, ptr(min_dist),
                                _stream()), "fsb_hull_min_dist")
    mask = torch.empty((N,), dtype=torch.uint8, device=dev)
    check(lib.fsb_hull_prune_mask(N, ptr(min_dist), lo, hi, ptr(_u8(add_mask)), ptr(mask), _stream()),
          "fsb_hull_prune_mask")
    return mask.bool()


@torch.no_grad()
def hull_pruning(model, optimizers, step: int, visual_hull: Tensor, scale_factor: float):
    """dn_model.py:1249-1276: prune Gaussians that float between 0.005 s and 0.02 s off the visual hull."""
    assert step == model.step
    if model.step <= model.config.warmup_length:
        return None
    hull_mask = hull_prune_mask(model.gauss_params["means"], visual_hull.to(model.gauss_params["means"].device),
                                scale_factor, getattr(model, "add_mask", None))
    model.max_2Dsize = None
    return _cull_only(model, optimizers, hull_mask)
