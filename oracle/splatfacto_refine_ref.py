"""TEST INFRASTRUCTURE — CPU/torch restatement of the densify / prune bookkeeping of the DN-Splatter model.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may import this module; the product path
(fusionsense_b200/densify.py) runs libfsb200's kernels and never falls back to it.

What it follows
  * refinement_after ........ /root/reference/dn_splatter/dn_model.py:326-451 (restated literally)
  * hull_pruning mask ....... /root/reference/dn_splatter/dn_model.py:1254-1269
  * remove_from_all_optim / dup_in_all_optim ... /root/reference/dn_splatter/dn_model.py:149-170
  * split_gaussians / dup_gaussians / cull_gaussians / dup_in_optim / remove_from_optim: nerfstudio==1.1.3
    `models/splatfacto.py` — NOT vendored in /root/reference (pyproject.toml:7) and not installed here; restated
    from the published release as summarised in SURVEY.md Appendix A.7.  **Parity unpinned** for these five
    (no reference test or golden vector exists for them); the dn_model.py driver above them is first-party
    code and is followed line by line.

The random samples of split_gaussians are an explicit argument so that the CUDA path and this restatement can be
fed the same draws (the reference calls torch.randn on the model's device).
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Dict, Optional

import torch
from torch import Tensor

PARAM_NAMES = ("means", "scales", "quats", "features_dc", "features_rest", "opacities", "normals")


@dataclass
class RefineConfig:
    """nerfstudio 1.1.3 SplatfactoModelConfig defaults with DN-Splatter's overrides (dn_model.py:113-133)."""
    warmup_length: int = 500
    refine_every: int = 100
    reset_alpha_every: int = 30
    stop_split_at: int = 15000
    densify_grad_thresh: float = 0.0008
    densify_size_thresh: float = 0.01
    n_split_samples: int = 2
    cull_alpha_thresh: float = 0.1
    cull_scale_thresh: float = 0.5
    cull_screen_size: float = 0.15
    split_screen_size: float = 0.05
    stop_screen_size_at: int = 4000
    continue_cull_post_densification: bool = True


def quat_to_rotmat(q: Tensor) -> Tensor:
    w, x, y, z = torch.unbind(q / q.norm(dim=-1, keepdim=True), dim=-1)
    return torch.stack([
        1 - 2 * (y * y + z * z), 2 * (x * y - w * z), 2 * (x * z + w * y),
        2 * (x * y + w * z), 1 - 2 * (x * x + z * z), 2 * (y * z - w * x),
        2 * (x * z - w * y), 2 * (y * z + w * x), 1 - 2 * (x * x + y * y)], dim=-1).reshape(q.shape[:-1] + (3, 3))


class RefState:
    """The slice of DNSplatterModel that refinement touches."""

    def __init__(self, params: Dict[str, Tensor], optim_state: Dict[str, Dict[str, Tensor]], config: RefineConfig,
                 step: int, num_train_data: int, last_size=(480, 640), add_mask: Optional[Tensor] = None):
        self.gauss_params = {k: v.clone() for k, v in params.items()}
        # optim_state[name] = {"exp_avg": ..., "exp_avg_sq": ...}  (one param per optimizer, dn_config.py:36-75)
        self.optim_state = {k: {m: t.clone() for m, t in v.items()} for k, v in optim_state.items()}
        self.config, self.step, self.num_train_data, self.last_size = config, step, num_train_data, last_size
        self.add_mask = add_mask
        self.xys_grad_norm = self.vis_counts = self.max_2Dsize = None

    # ---- nerfstudio splatfacto (A.7) ---------------------------------------------------------
    def split_gaussians(self, split_mask: Tensor, samps: int, samples: Tensor):
        n_splits = int(split_mask.sum())
        gp = self.gauss_params
        centered_samples = samples.reshape(samps * n_splits, 3)
        scaled_samples = torch.exp(gp["scales"][split_mask].repeat(samps, 1)) * centered_samples
        quats = gp["quats"][split_mask] / gp["quats"][split_mask].norm(dim=-1, keepdim=True)
        rots = quat_to_rotmat(quats.repeat(samps, 1))
        rotated_samples = torch.bmm(rots, scaled_samples[..., None]).squeeze(-1)
        new_means = rotated_samples + gp["means"][split_mask].repeat(samps, 1)
        size_fac = 1.6
        new_scales = torch.log(torch.exp(gp["scales"][split_mask]) / size_fac).repeat(samps, 1)
        gp["scales"][split_mask] = torch.log(torch.exp(gp["scales"][split_mask]) / size_fac)
        out = {"means": new_means, "scales": new_scales}
        for name, param in gp.items():
            if name not in out:
                out[name] = param[split_mask].repeat(samps, *([1] * (param.dim() - 1)))
        return out

    def dup_gaussians(self, dup_mask: Tensor):
        return {name: param[dup_mask] for name, param in self.gauss_params.items()}

    def cull_gaussians(self, extra_cull_mask: Optional[Tensor] = None) -> Tensor:
        cfg, gp = self.config, self.gauss_params
        culls = (torch.sigmoid(gp["opacities"]) < cfg.cull_alpha_thresh).squeeze()
        if extra_cull_mask is not None:
            culls = culls | extra_cull_mask
        if self.step > cfg.refine_every * cfg.reset_alpha_every:
            toobigs = (torch.exp(gp["scales"]).max(dim=-1).values > cfg.cull_scale_thresh).squeeze()
            if self.step < cfg.stop_screen_size_at:
                if self.max_2Dsize is not None:
                    toobigs = toobigs | (self.max_2Dsize > cfg.cull_screen_size).squeeze()
            culls = culls | toobigs
        for name, param in gp.items():
            gp[name] = param[~culls]
        return culls

    def dup_in_all_optim(self, dup_idcs: Tensor, n: int):
        for st in self.optim_state.values():
            for key in ("exp_avg", "exp_avg_sq"):
                t = st[key]
                reps = tuple(1 for _ in range(t.dim() - 1))
                st[key] = torch.cat([t, torch.zeros_like(t[dup_idcs]).repeat(n, *reps)], dim=0)
        if self.add_mask is not None:
            self.add_mask = torch.cat([self.add_mask, torch.zeros(dup_idcs.shape[0] * n, dtype=self.add_mask.dtype)])

    def remove_from_all_optim(self, deleted_mask: Tensor):
        for st in self.optim_state.values():
            for key in ("exp_avg", "exp_avg_sq"):
                st[key] = st[key][~deleted_mask]
        if self.add_mask is not None:
            self.add_mask = self.add_mask[~deleted_mask]

    # ---- dn_model.py:326-451 -----------------------------------------------------------------
    def refinement_after(self, samples: Optional[Tensor] = None):
        cfg = self.config
        if self.step <= cfg.warmup_length:
            return None
        gp = self.gauss_params
        reset_interval = cfg.reset_alpha_every * cfg.refine_every
        do_densification = (self.step < cfg.stop_split_at
                            and self.step % reset_interval > self.num_train_data + cfg.refine_every)
        deleted_mask = None
        if do_densification:
            avg_grad_norm = (self.xys_grad_norm / self.vis_counts) * 0.5 * max(self.last_size[0], self.last_size[1])
            high_grads = (avg_grad_norm > cfg.densify_grad_thresh).squeeze()
            splits = (gp["scales"].exp().max(dim=-1).values > cfg.densify_size_thresh).squeeze()
            if self.step < cfg.stop_screen_size_at:
                splits |= (self.max_2Dsize > cfg.split_screen_size).squeeze()
            splits &= high_grads
            if self.add_mask is not None:
                splits &= ~self.add_mask
            nsamps = cfg.n_split_samples
            # order matters (dn_model.py:369-375): split_gaussians shrinks the split parents' scales IN PLACE before
            # `dups` is built from the scales, so a split parent (high gradient by construction) whose shrunk scale
            # falls to <= densify_size_thresh is also duplicated, and that copy survives the parent's cull.
            split_params = self.split_gaussians(splits, nsamps, samples)
            dups = (gp["scales"].exp().max(dim=-1).values <= cfg.densify_size_thresh).squeeze()
            dups &= high_grads
            if self.add_mask is not None:
                dups &= ~self.add_mask
            dup_params = self.dup_gaussians(dups)
            for name, param in gp.items():
                gp[name] = torch.cat([param, split_params[name], dup_params[name]], dim=0)
            self.max_2Dsize = torch.cat([self.max_2Dsize, torch.zeros_like(split_params["scales"][:, 0]),
                                         torch.zeros_like(dup_params["scales"][:, 0])], dim=0)
            split_idcs = torch.where(splits)[0]
            self.dup_in_all_optim(split_idcs, nsamps)
            dup_idcs = torch.where(dups)[0]
            self.dup_in_all_optim(dup_idcs, 1)
            splits_mask = torch.cat((splits, torch.zeros(nsamps * int(splits.sum()) + int(dups.sum()), dtype=torch.bool)))
            deleted_mask = self.cull_gaussians(splits_mask)
        elif self.step >= cfg.stop_split_at and cfg.continue_cull_post_densification:
            deleted_mask = self.cull_gaussians()
        if deleted_mask is not None:
            self.remove_from_all_optim(deleted_mask)
        if self.step < cfg.stop_split_at and self.step % reset_interval == cfg.refine_every:
            reset_value = cfg.cull_alpha_thresh * 2.0
            gp["opacities"] = torch.clamp(gp["opacities"], max=torch.logit(torch.tensor(reset_value)).item())
            st = self.optim_state["opacities"]
            st["exp_avg"] = torch.zeros_like(st["exp_avg"])
            st["exp_avg_sq"] = torch.zeros_like(st["exp_avg_sq"])
        self.xys_grad_norm = self.vis_counts = self.max_2Dsize = None
        return deleted_mask

    # ---- dn_model.py:1249-1276 ---------------------------------------------------------------
    def hull_mask(self, visual_hull: Tensor, scale_factor: float) -> Tensor:
        means = self.gauss_params["means"]
        center = visual_hull.mean(dim=0)
        close_mask = torch.norm(means - center, dim=1) <= 0.2 * scale_factor
        filtered_means = means[close_mask]
        # brute force in the input precision (torch.cdist's matmul shortcut is less accurate than this)
        distances = (filtered_means[:, None, :] - visual_hull[None, :, :]).norm(dim=-1)
        min_distances = distances.min(dim=-1).values
        filtered = (min_distances > 0.005 * scale_factor) & (min_distances <= 0.02 * scale_factor)
        hull_mask = torch.zeros(means.shape[0], dtype=torch.bool)
        hull_mask[close_mask] = filtered
        if self.add_mask is not None:
            hull_mask[self.add_mask] = False
        return hull_mask

    def hull_pruning(self, visual_hull: Tensor, scale_factor: float):
        if self.step <= self.config.warmup_length:
            return None
        hull_mask = self.hull_mask(visual_hull, scale_factor)
        self.max_2Dsize = None
        deleted_mask = self.cull_gaussians(hull_mask)
        self.remove_from_all_optim(deleted_mask)
        return deleted_mask
