"""TEST INFRASTRUCTURE — plain-torch restatement of the level-set search inside
`DNSplatterModel.compute_level_surface_points` (dn_splatter/dn_model.py:1766-1880): std of the first neighbour, 21 ray
samples, densities (the inlined get_density: normalised above 1, not clamped), first crossing per surface level.

Pinned: tests/golden/level_set.npz holds the outputs of the reference's own function (the unbound
`DNSplatterModel.compute_level_surface_points` run on the CPU with a stand-in `self`, oracle/make_golden_level_set.py);
tests/test_level_set.py checks this restatement against them.  Only tests/ and tools/ may import this module.
"""
from __future__ import annotations

import torch

from oracle.knn_ref import quat_to_rotmat_ref


def level_crossings_ref(points, cam_pos, closest, means, log_scales, quats, opacities, surface_levels):
    """-> (t [L,P], valid [L,P] bool, densities [P,21], std [P]); every statement in the reference's order."""
    viewdirs = -means + cam_pos[None, :]
    viewdirs = viewdirs / viewdirs.norm(dim=-1, keepdim=True)
    q = quats / quats.norm(dim=-1, keepdim=True)
    inv_rots = quat_to_rotmat_ref(q * torch.tensor([1.0, -1.0, -1.0, -1.0]))
    stds = (torch.exp(log_scales) * torch.bmm(inv_rots, viewdirs[..., None])[..., 0]).norm(dim=-1)
    points_stds = stds[closest][..., 0]
    n_in_range = 21
    points_range = torch.linspace(-3, 3, n_in_range).view(1, -1, 1) * points_stds[..., None, None].expand(-1, n_in_range, 1)
    camera_to_samples = torch.nn.functional.normalize(points - cam_pos[None, :], dim=-1)
    samples = (points[:, None, :] + points_range * camera_to_samples[:, None, :]).view(-1, 3)
    K = closest.shape[1]
    sidx = closest[:, None, :].expand(-1, n_in_range, -1).reshape(-1, K)
    strengths = torch.sigmoid(opacities)
    scale = 1.0 / torch.exp(log_scales).clamp(min=1e-3)
    M = quat_to_rotmat_ref(quats) * scale[..., None, :]
    shift = samples[:, None] - means[sidx]
    man = M[sidx].transpose(-1, -2) @ shift[..., None]
    nb = (man[..., 0] * man[..., 0]).sum(dim=-1).clamp(min=0.0, max=1e8)
    nb = strengths[sidx][..., 0] * torch.exp(-0.5 * nb)
    dens = nb.sum(dim=-1)
    big = dens >= 1.0
    dens[big] = dens[big] / (dens[big] + 1e-5)
    dens = dens.reshape(-1, n_in_range)
    ts, valids = [], []
    for level in surface_levels:
        under = dens - level < 0
        above = dens - level > 0
        _, first = above.max(dim=-1, keepdim=True)
        empty = ~under[..., 0] + (first[..., 0] == 0)
        safe = first.clamp(min=1)
        v1 = dens.gather(-1, safe).view(-1)
        v0 = dens.gather(-1, safe - 1).view(-1)
        rng = points_range[..., 0]
        t1 = rng.gather(-1, safe).view(-1)
        t0 = rng.gather(-1, safe - 1).view(-1)
        t = (level - v0) / (v1 - v0) * (t1 - t0) + t0
        ts.append(torch.where(empty, torch.zeros_like(t), t))
        valids.append(~empty)
    return torch.stack(ts), torch.stack(valids), dens, points_stds
