"""Drop-in for `gsplat.rendering` (gsplat==1.0.0) backed by libfsb200's sm_100a kernels.

Only what FusionSense / DN-Splatter and nerfstudio 1.1.3's splatfacto reach is provided:
`rasterization()` with the exact gsplat 1.0.0 signature, defaults, return tuple and meta keys
(reference call site: /root/reference/dn_splatter/dn_model.py:570-591; semantics restated in
SURVEY.md Appendix A.1).
"""
from __future__ import annotations

import math
from typing import Dict, Optional, Tuple

import torch
from torch import Tensor
from typing_extensions import Literal

from .. import ops
from .cuda_legacy._wrapper import remember_binning


def rasterization(
    means: Tensor,  # [N, 3]
    quats: Tensor,  # [N, 4]
    scales: Tensor,  # [N, 3]
    opacities: Tensor,  # [N]
    colors: Tensor,  # [(C,) N, D] or [(C,) N, K, 3]
    viewmats: Tensor,  # [C, 4, 4]
    Ks: Tensor,  # [C, 3, 3]
    width: int,
    height: int,
    near_plane: float = 0.01,
    far_plane: float = 1e10,
    radius_clip: float = 0.0,
    eps2d: float = 0.3,
    sh_degree: Optional[int] = None,
    packed: bool = True,
    tile_size: int = 16,
    backgrounds: Optional[Tensor] = None,
    render_mode: Literal["RGB", "D", "ED", "RGB+D", "RGB+ED"] = "RGB",
    sparse_grad: bool = False,
    absgrad: bool = False,
    rasterize_mode: Literal["classic", "antialiased"] = "classic",
    channel_chunk: int = 32,
) -> Tuple[Tensor, Tensor, Dict]:
    """Rasterize 3D Gaussians to `[C, H, W, D]` images; see gsplat 1.0.0 `rasterization` for the contract.

    Differences from upstream, all deliberate: (1) culled Gaussians have zeroed (not uninitialised)
    `means2d / depths / conics`; (2) `packed=True` and `sparse_grad=True` change gsplat's meta layout to a
    flattened nnz form that neither DN-Splatter nor splatfacto use — they are refused loudly instead of being
    emulated; (3) there is no CPU path.
    """
    return _rasterization_impl(means, quats, scales, opacities, colors, viewmats, Ks, width, height, near_plane,
                               far_plane, radius_clip, eps2d, sh_degree, packed, tile_size, backgrounds, render_mode,
                               sparse_grad, absgrad, rasterize_mode, channel_chunk)


def rasterization_from_params(
    means: Tensor,  # [N, 3]
    quats: Tensor,  # [N, 4]  as stored (any norm)
    log_scales: Tensor,  # [N, 3]  as stored (exp applied inside the projection kernel)
    opacities: Tensor,  # [N]  activated (sigmoid applied by the caller, shared with the normals pass)
    features_dc: Tensor,  # [N, 3]
    features_rest: Tensor,  # [N, K-1, 3]
    viewmats: Tensor,
    Ks: Tensor,
    width: int,
    height: int,
    sh_degree: int,
    near_plane: float = 0.01,
    far_plane: float = 1e10,
    radius_clip: float = 0.0,
    eps2d: float = 0.3,
    tile_size: int = 16,
    backgrounds: Optional[Tensor] = None,
    render_mode: Literal["RGB", "RGB+D", "RGB+ED"] = "RGB",
    absgrad: bool = False,
    prune_lists: bool = False,
    colors_b: Optional[Tensor] = None,  # [N, 3]
    backgrounds_b: Optional[Tensor] = None,  # [C, 3]
) -> Tuple[Tensor, Tensor, Dict]:
    """Extension (not part of gsplat): `rasterization(packed=False, rasterize_mode="classic")` of
    `quats / |quats|, exp(log_scales), cat(features_dc[:, None], features_rest)` with those three torch
    expressions (dn_model.py:566-574) evaluated inside the projection kernel and differentiated in place.
    Same return tuple and meta keys as `rasterization()`.

    `prune_lists`: bin only the (Gaussian, tile) pairs that can pass the alpha test somewhere in the tile
    (csrc/isect_reach.cu).  Images and gradients are unchanged; `meta["isect_ids"] / ["flatten_ids"] /
    ["isect_offsets"]` then hold the pruned lists instead of gsplat's bounding-box lists, `tiles_per_gauss` stays
    the bounding-box count.

    `colors_b` (+ `backgrounds_b`): a second colour set composited by the SAME walk with the semantics of
    `gsplat.rasterize_gaussians(means2d.detach(), depths, radii, conics, tiles_per_gauss, colors_b, opacities[:, None],
    H, W, tile_size, background=backgrounds_b)` — the normals pass of dn_model.py:644-653 — including the 0.1.x
    bounding-box rule of that call (union lists with FSB_LEGACY_FLAG, include/fsb200.h).  The result comes back as
    `meta["render_b"]` [C, H, W, 3]; implies `prune_lists`."""
    assert render_mode in ["RGB", "RGB+D", "RGB+ED"], render_mode
    return _rasterization_impl(means, quats, log_scales, opacities, None, viewmats, Ks, width, height, near_plane,
                               far_plane, radius_clip, eps2d, sh_degree, False, tile_size, backgrounds, render_mode,
                               False, absgrad, "classic", 32, stored=(features_dc, features_rest),
                               prune=prune_lists or colors_b is not None, colors_b=colors_b, backgrounds_b=backgrounds_b)


class _Meta(dict):
    """meta dict whose "isect_ids" is built on first access when the lists came from the two-level binning
    (ops.PackedIsectIds): the render path itself never needs the gsplat key form."""

    def __getitem__(self, key):
        v = dict.__getitem__(self, key)
        if isinstance(v, ops.PackedIsectIds):
            v = v.materialize()
            dict.__setitem__(self, key, v)
        return v

    def get(self, key, default=None):
        return self[key] if key in self else default


def _rasterization_impl(means, quats, scales, opacities, colors, viewmats, Ks, width, height, near_plane, far_plane,
                        radius_clip, eps2d, sh_degree, packed, tile_size, backgrounds, render_mode, sparse_grad,
                        absgrad, rasterize_mode, channel_chunk, stored=None, prune=False, colors_b=None,
                        backgrounds_b=None):
    N = means.shape[0]
    C = viewmats.shape[0]
    assert means.shape == (N, 3), means.shape
    assert quats.shape == (N, 4), quats.shape
    assert scales.shape == (N, 3), scales.shape
    assert opacities.shape == (N,), opacities.shape
    assert viewmats.shape == (C, 4, 4), viewmats.shape
    assert Ks.shape == (C, 3, 3), Ks.shape
    assert render_mode in ["RGB", "D", "ED", "RGB+D", "RGB+ED"], render_mode
    assert rasterize_mode in ["classic", "antialiased"], rasterize_mode
    if not means.is_cuda:
        raise RuntimeError("fusionsense_b200.gsplat.rasterization needs CUDA tensors (no CPU fallback)")
    if packed or sparse_grad:
        raise NotImplementedError(
            "fusionsense_b200's rasterization implements the unpacked layout (packed=False, sparse_grad=False) that "
            "DN-Splatter (dn_model.py:580,586) and splatfacto use"
        )

    if stored is not None:
        assert sh_degree is not None and stored[0].shape == (N, 3), stored[0].shape
    elif sh_degree is None:
        assert (colors.dim() == 2 and colors.shape[0] == N) or (
            colors.dim() == 3 and colors.shape[:2] == (C, N)
        ), colors.shape
    else:
        assert (colors.dim() == 3 and colors.shape[0] == N and colors.shape[2] == 3) or (
            colors.dim() == 4 and colors.shape[:2] == (C, N) and colors.shape[3] == 3
        ), colors.shape
        assert (sh_degree + 1) ** 2 <= colors.shape[-2], colors.shape
        if colors.dim() == 4:
            raise NotImplementedError("per-camera SH coefficients [C,N,K,3] are not used by the reference")

    want_depth = render_mode in ["RGB+D", "RGB+ED", "D", "ED"]
    only_depth = render_mode in ["D", "ED"]
    ed_normalize = render_mode in ["ED", "RGB+ED"]
    use_sh = sh_degree is not None and not only_depth
    calc_comp = rasterize_mode == "antialiased"

    # colour evaluation fused with the projection when SH is on
    if use_sh:
        # camera centres: evaluated inside the kernel (-R^-1 t) unless a gradient has to reach the view matrices
        campos = torch.linalg.inv(viewmats)[:, :3, 3] if viewmats.requires_grad else None
        color_stride = 4 if want_depth else 3
        depth_channel = 3 if want_depth else -1
        coeffs = colors
    else:
        campos, coeffs, color_stride, depth_channel = None, None, 0, -1

    # totals[0] <- n_isects (scan), totals[1] <- tiles by which the legacy 0.1.x bbox rule would differ; one D2H read
    totals = torch.zeros(2, dtype=torch.int64, device=means.device)
    if stored is not None:
        comps = None
        radii, means2d, depths, conics, sh_colors, tiles_per_gauss = ops.ProjectParams.apply(
            means, quats, scales, stored[0], stored[1], viewmats, Ks, width, height, eps2d, near_plane, far_plane,
            radius_clip, tile_size, sh_degree, color_stride, depth_channel, totals[1:] if C == 1 else None)
    else:
        radii, means2d, depths, conics, comps, sh_colors, tiles_per_gauss = ops.ProjectSH.apply(
            means, quats, scales, coeffs, viewmats, Ks, campos, width, height, eps2d, near_plane, far_plane,
            radius_clip, tile_size, sh_degree if use_sh else None, color_stride, depth_channel, calc_comp,
            totals[1:] if C == 1 else None)

    # static-capacity mode (captured step): a caller may start work that needs only the projection outputs (the
    # DN-Splatter normals pass) on another stream while this one bins and composites
    proj_done = None
    if ops.static_mode() is not None:
        proj_done = torch.cuda.Event()
        proj_done.record()

    opac = opacities[None].expand(C, N)
    if comps is not None:
        opac = opac * comps
    opac = opac.contiguous()

    tile_width = math.ceil(width / float(tile_size))
    tile_height = math.ceil(height / float(tile_size))
    with torch.no_grad():
        _, isect_ids, flatten_ids, isect_offsets = ops.isect_tiles(
            means2d, radii, depths, tile_size, tile_width, tile_height, tiles_per_gauss=tiles_per_gauss,
            totals=totals, reach=(conics, opac) if prune else None, legacy_bbox=2 if colors_b is not None else False,
            lazy_ids=True)
        n_dev = getattr(flatten_ids, "n_dev", None)  # static-capacity mode (ops.static_capacity): count on device
        lists_done = None
        if n_dev is not None:
            lists_done = torch.cuda.Event()
            lists_done.record()
        remember_binning(means2d, depths, radii, width, height, tile_size, flatten_ids.numel(), flatten_ids,
                         isect_offsets, legacy_extra=totals.host[1] if (C == 1 and n_dev is None) else None,
                         lists_done=lists_done, pruned=prune, union=colors_b is not None)

    if use_sh:
        ras_colors = sh_colors  # [C, N, 3 or 4], depth already in channel 3
    elif only_depth:
        ras_colors = depths[..., None]
    else:
        ras_colors = colors if colors.dim() == 3 else colors[None].expand(C, N, colors.shape[-1])
        if want_depth:
            ras_colors = torch.cat([ras_colors, depths[..., None]], dim=-1)
    ras_colors = ras_colors.contiguous()

    D = ras_colors.shape[-1]
    if backgrounds is not None:
        assert backgrounds.shape == (C, D), backgrounds.shape

    def _raster(cols, bgs, ed):
        d = cols.shape[-1]
        dp = ops.supported_channels(d)
        if dp != d:
            cols = torch.cat([cols, cols.new_zeros(*cols.shape[:-1], dp - d)], dim=-1)
            if bgs is not None:
                bgs = torch.cat([bgs, bgs.new_zeros(C, dp - d)], dim=-1)
            # the ED channel must stay last for the fused normalisation; fall back to doing it outside
            fused_ed = False
        else:
            fused_ed = ed
        out, alpha = ops.RasterizeToPixels.apply(means2d, conics, cols, opac, bgs, None, width, height, tile_size,
                                                 isect_offsets, flatten_ids, absgrad, fused_ed, n_dev)
        if dp != d:
            out = out[..., :d]
        if ed and not fused_ed:
            out = torch.cat([out[..., :-1], out[..., -1:] / alpha.clamp(min=1e-10)], dim=-1)
        return out, alpha

    render_b = None
    if colors_b is not None:
        # both colour sets in one walk (csrc/raster.cu); set A must be the 4-channel RGB + depth colours
        assert D == 4 and colors_b.shape[-1] == 3, (D, colors_b.shape)
        cb = colors_b if colors_b.dim() == 3 else colors_b[None].expand(C, N, 3)
        if backgrounds_b is None:
            backgrounds_b = torch.ones((C, 3), dtype=torch.float32, device=means.device)
        render_colors, render_b, render_alphas = ops.RasterizeDN.apply(
            means2d, conics, ras_colors, cb.contiguous(), opac, backgrounds, backgrounds_b, width, height, tile_size,
            isect_offsets, flatten_ids, absgrad, 3 if ed_normalize else -1, n_dev)
    elif D > channel_chunk:
        n_chunks = (D + channel_chunk - 1) // channel_chunk
        outs, render_alphas = [], None
        for i in range(n_chunks):
            lo, hi = i * channel_chunk, min((i + 1) * channel_chunk, D)
            bgs = backgrounds[..., lo:hi] if backgrounds is not None else None
            o, a = _raster(ras_colors[..., lo:hi].contiguous(), bgs, ed_normalize and hi == D)
            outs.append(o)
            render_alphas = a
        render_colors = torch.cat(outs, dim=-1)
    else:
        render_colors, render_alphas = _raster(ras_colors, backgrounds, ed_normalize)

    meta = _Meta({
        "camera_ids": None,
        "gaussian_ids": None,
        "radii": radii,
        "means2d": means2d,
        "depths": depths,
        "conics": conics,
        "opacities": opac,
        "tile_width": tile_width,
        "tile_height": tile_height,
        "tiles_per_gauss": tiles_per_gauss,
        "isect_ids": isect_ids,
        "flatten_ids": flatten_ids,
        "isect_offsets": isect_offsets,
        "width": width,
        "height": height,
        "tile_size": tile_size,
        "n_cameras": C,
    })
    if render_b is not None:
        meta["render_b"] = render_b  # extension: the second colour set (rasterization_from_params(colors_b=...))
    if proj_done is not None:
        meta["projection_done"] = proj_done  # extension, static-capacity mode only
    return render_colors, render_alphas, meta
