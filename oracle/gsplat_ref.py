"""CPU oracle for the render path: a plain-torch restatement of what FusionSense reaches in gsplat==1.0.0.

TEST INFRASTRUCTURE ONLY.  Nothing under oracle/ is imported by the product package
(fusionsense_b200/); only tests/, __graft_entry__.smoke() and bench.py's CPU-baseline legs use it.

PARITY UNPINNED: the algorithm lives in the third-party dependency `gsplat==1.0.0`
(/root/reference/pyproject.toml:8, env1.yml:283), which is neither vendored under /root/reference nor
installed in this image, and the reference ships no tests or golden vectors for this path
(SURVEY.md §4, §8c).  This file restates the published gsplat 1.0.0 algorithm
and anchors on the reference's own call sites:
  * /root/reference/dn_splatter/dn_model.py:570-591   gsplat.rendering.rasterization(...)
  * /root/reference/dn_splatter/dn_model.py:644-653   gsplat.rasterize_gaussians(...)
  * /root/reference/dn_splatter/dn_model.py:286,623   gsplat.cuda_legacy._torch_impl.quat_to_rotmat
Self-consistency is pinned instead: the analytic structure is checked against fp64 autograd and brute-force
per-pixel loops in tests/test_oracle_gsplat.py.

Everything is differentiable torch (any float dtype, CPU); integer stages are exact.
"""
from __future__ import annotations

import math
from typing import Dict, Optional, Tuple

import torch
from torch import Tensor

ALPHA_MAX = 0.999
ALPHA_MIN = 1.0 / 255.0
T_MIN = 1e-4


# --------------------------------------------------------------------------------------------
# A.2 projection  (gsplat/cuda/_torch_impl.py::_fully_fused_projection and friends)
# --------------------------------------------------------------------------------------------
def quat_to_rotmat(quats: Tensor) -> Tensor:
    """[..., 4] wxyz -> [..., 3, 3]; normalises internally (gsplat `_quat_to_rotmat`)."""
    q = quats / quats.norm(dim=-1, keepdim=True)
    w, x, y, z = q.unbind(-1)
    R = torch.stack(
        [
            1 - 2 * (y * y + z * z), 2 * (x * y - w * z), 2 * (x * z + w * y),
            2 * (x * y + w * z), 1 - 2 * (x * x + z * z), 2 * (y * z - w * x),
            2 * (x * z - w * y), 2 * (y * z + w * x), 1 - 2 * (x * x + y * y),
        ],
        dim=-1,
    )
    return R.reshape(quats.shape[:-1] + (3, 3))


def world_to_cam_means(means: Tensor, viewmats: Tensor) -> Tensor:
    """mean_c = R mean + t with every product and sum rounded separately, left to right.

    The camera-space z is the low half of the 64-bit sort key, so kernel and oracle must agree bit for bit:
    both compute ((R0*x + R1*y) + R2*z) + t without FMA contraction.
    """
    R = viewmats[:, :3, :3]  # [C,3,3]
    t = viewmats[:, :3, 3]  # [C,3]
    x, y, z = means[:, 0][None, :, None], means[:, 1][None, :, None], means[:, 2][None, :, None]
    a = R[:, None, :, 0] * x
    b = R[:, None, :, 1] * y
    c = R[:, None, :, 2] * z
    return ((a + b) + c) + t[:, None, :]  # [C,N,3]


def fully_fused_projection(means, quats, scales, viewmats, Ks, width, height, eps2d=0.3, near_plane=0.01,
                           far_plane=1e10, radius_clip=0.0, calc_compensations=False):
    """-> radii [C,N] int32, means2d [C,N,2], depths [C,N], conics [C,N,3], compensations [C,N] | None.

    Culled entries are zero (the kernels do the same; gsplat leaves them uninitialised).
    """
    C, N = viewmats.shape[0], means.shape[0]
    dt = means.dtype
    Rq = quat_to_rotmat(quats)  # [N,3,3]
    M = Rq * scales[:, None, :]
    Sigma = M @ M.transpose(-1, -2)  # [N,3,3]
    Rv = viewmats[:, :3, :3]
    mc = world_to_cam_means(means, viewmats)  # [C,N,3]
    SigmaC = Rv[:, None] @ Sigma[None] @ Rv[:, None].transpose(-1, -2)  # [C,N,3,3]

    fx, fy = Ks[:, 0, 0][:, None], Ks[:, 1, 1][:, None]
    cx, cy = Ks[:, 0, 2][:, None], Ks[:, 1, 2][:, None]
    x, y, z = mc.unbind(-1)
    zs = torch.where(z.abs() < 1e-30, torch.full_like(z, 1e-30), z)  # keep culled lanes finite
    tan_fovx = 0.5 * width / fx
    tan_fovy = 0.5 * height / fy
    lim_x, lim_y = 1.3 * tan_fovx, 1.3 * tan_fovy
    tx = zs * torch.minimum(lim_x, torch.maximum(-lim_x, x / zs))
    ty = zs * torch.minimum(lim_y, torch.maximum(-lim_y, y / zs))
    O = torch.zeros_like(z)
    J = torch.stack([fx / zs, O, -fx * tx / zs**2, O, fy / zs, -fy * ty / zs**2], dim=-1).reshape(C, N, 2, 3)
    cov2d = J @ SigmaC @ J.transpose(-1, -2)  # [C,N,2,2]
    means2d = torch.stack([fx * x / zs + cx, fy * y / zs + cy], dim=-1)

    c_xx, c_xy, c_yy = cov2d[..., 0, 0], cov2d[..., 0, 1], cov2d[..., 1, 1]
    det0 = c_xx * c_yy - c_xy * c_xy
    b_xx, b_yy = c_xx + eps2d, c_yy + eps2d
    det = b_xx * b_yy - c_xy * c_xy
    valid = (z >= near_plane) & (z <= far_plane) & (det > 0)
    det_s = torch.where(det > 0, det, torch.ones_like(det))
    conics = torch.stack([b_yy / det_s, -c_xy / det_s, b_xx / det_s], dim=-1)
    comp = torch.sqrt(torch.clamp(det0 / det_s, min=0.0))
    b = 0.5 * (b_xx + b_yy)
    v1 = b + torch.sqrt(torch.clamp(b * b - det_s, min=0.01))
    radius = torch.ceil(3.0 * torch.sqrt(v1)).detach()
    valid = valid & (radius > radius_clip)
    inside = ((means2d[..., 0] + radius > 0) & (means2d[..., 0] - radius < width)
              & (means2d[..., 1] + radius > 0) & (means2d[..., 1] - radius < height))
    valid = valid & inside
    radii = torch.where(valid, radius, torch.zeros_like(radius)).to(torch.int32)
    vf = valid.to(dt)
    means2d = means2d * vf[..., None]
    depths = z * vf
    conics = conics * vf[..., None]
    comp = comp * vf if calc_compensations else None
    return radii, means2d, depths, conics, comp


# --------------------------------------------------------------------------------------------
# A.3 spherical harmonics  (gsplat `_spherical_harmonics`, Sloan fast form)
# --------------------------------------------------------------------------------------------
def sh_bases(degree: int, dirs: Tensor) -> Tensor:
    """dirs [...,3] (normalised inside) -> [..., (degree+1)^2]."""
    d = dirs / dirs.norm(dim=-1, keepdim=True)
    x, y, z = d.unbind(-1)
    out = [torch.full_like(x, 0.2820947917738781)]
    if degree >= 1:
        out += [-0.48860251190292 * y, 0.48860251190292 * z, -0.48860251190292 * x]
    if degree >= 2:
        z2 = z * z
        fTmp0B = -1.092548430592079 * z
        fC1 = x * x - y * y
        fS1 = 2 * x * y
        out += [0.5462742152960395 * fS1, fTmp0B * y, 0.9461746957575601 * z2 - 0.3153915652525201, fTmp0B * x,
                0.5462742152960395 * fC1]
    if degree >= 3:
        fTmp0C = -2.285228997322329 * z2 + 0.4570457994644658
        fTmp1B = 1.445305721320277 * z
        fC2 = x * fC1 - y * fS1
        fS2 = x * fS1 + y * fC1
        out += [-0.5900435899266435 * fS2, fTmp1B * fS1, fTmp0C * y,
                z * (1.865881662950577 * z2 - 1.119528997770346), fTmp0C * x, fTmp1B * fC1,
                -0.5900435899266435 * fC2]
    if degree >= 4:
        raise NotImplementedError("degree <= 3 (dn_model.py:562-565)")
    return torch.stack(out, dim=-1)


def spherical_harmonics(degree: int, dirs: Tensor, coeffs: Tensor, masks: Optional[Tensor] = None) -> Tensor:
    """dirs [...,3], coeffs [...,K,3] -> [...,3]; masked-out entries are 0."""
    nb = (degree + 1) ** 2
    safe = dirs
    if masks is not None:
        safe = torch.where(masks[..., None], dirs, torch.ones_like(dirs))
    B = sh_bases(degree, safe)  # [..., nb]
    col = (B[..., :, None] * coeffs[..., :nb, :]).sum(dim=-2)
    if masks is not None:
        col = col * masks[..., None].to(col.dtype)
    return col


# --------------------------------------------------------------------------------------------
# A.4 tile intersection, keys, sort, offsets  (integer; numpy-exact through torch int ops)
# --------------------------------------------------------------------------------------------
def tile_bbox(means2d: Tensor, radii: Tensor, tile_size: int, tile_w: int, tile_h: int, legacy_bbox: bool = False):
    m = means2d.to(torch.float32)
    ts = torch.tensor(float(tile_size), dtype=torch.float32)
    tr = radii.to(torch.float32) / ts
    tx, ty = m[..., 0] / ts, m[..., 1] / ts
    if legacy_bbox:
        x0, y0 = torch.trunc(tx - tr), torch.trunc(ty - tr)
        x1, y1 = torch.trunc(tx + tr + 1.0), torch.trunc(ty + tr + 1.0)
    else:
        x0, y0 = torch.floor(tx - tr), torch.floor(ty - tr)
        x1, y1 = torch.ceil(tx + tr), torch.ceil(ty + tr)
    x0 = x0.to(torch.int64).clamp(0, tile_w)
    x1 = x1.to(torch.int64).clamp(0, tile_w)
    y0 = y0.to(torch.int64).clamp(0, tile_h)
    y1 = y1.to(torch.int64).clamp(0, tile_h)
    return x0, y0, x1, y1


def isect_tiles(means2d: Tensor, radii: Tensor, depths: Tensor, tile_size: int, tile_w: int, tile_h: int,
                sort: bool = True, legacy_bbox: bool = False):
    """-> tiles_per_gauss [C,N] int32, isect_ids [n_isects] int64, flatten_ids [n_isects] int32."""
    C, N = radii.shape
    x0, y0, x1, y1 = tile_bbox(means2d, radii, tile_size, tile_w, tile_h, legacy_bbox)
    vis = radii > 0
    cnt = ((y1 - y0) * (x1 - x0)) * vis
    tiles_per_gauss = cnt.to(torch.int32)
    flat_cnt = cnt.reshape(-1)
    n_isects = int(flat_cnt.sum())
    tile_bits = int(tile_w * tile_h).bit_length()
    # expand every (c, n) into its cnt entries, row-major over (i in [y0,y1), j in [x0,x1))
    owner = torch.repeat_interleave(torch.arange(C * N), flat_cnt)
    start = torch.cumsum(flat_cnt, 0) - flat_cnt
    local = torch.arange(n_isects) - start[owner]
    w = (x1 - x0).reshape(-1)[owner].clamp(min=1)
    ti = y0.reshape(-1)[owner] + local // w
    tj = x0.reshape(-1)[owner] + local % w
    tile_id = ti * tile_w + tj
    cam = owner // N
    depth_bits = depths.to(torch.float32).contiguous().view(torch.int32).reshape(-1)[owner].to(torch.int64)
    isect_ids = (cam << (32 + tile_bits)) | (tile_id << 32) | (depth_bits & 0xFFFFFFFF)
    flatten_ids = owner.to(torch.int32)
    if sort:
        order = torch.argsort(isect_ids, stable=True)
        isect_ids, flatten_ids = isect_ids[order], flatten_ids[order]
    return tiles_per_gauss, isect_ids, flatten_ids


def isect_offset_encode(isect_ids: Tensor, C: int, tile_w: int, tile_h: int) -> Tensor:
    n_tiles = tile_w * tile_h
    tile_bits = int(n_tiles).bit_length()
    k = isect_ids >> 32
    lin = (k >> tile_bits) * n_tiles + (k & ((1 << tile_bits) - 1))
    offsets = torch.searchsorted(lin, torch.arange(C * n_tiles), right=False)
    return offsets.to(torch.int32).reshape(C, tile_h, tile_w)


# --------------------------------------------------------------------------------------------
# A.5 compositing  (gsplat rasterize_to_pixels_{fwd,bwd}_kernel semantics; backward via autograd)
# --------------------------------------------------------------------------------------------
def rasterize_to_pixels(means2d, conics, colors, opacities, width, height, tile_size, isect_offsets, flatten_ids,
                        backgrounds=None, return_last_ids=False, stats=None):
    """means2d [C,N,2] conics [C,N,3] colors [C,N,D] opacities [C,N] -> colors [C,H,W,D], alphas [C,H,W,1].

    `stats` (optional dict) receives the pair counts of SURVEY.md §8d: "blended" = (pixel, entry) pairs composited,
    "visited" = entries a per-pixel walk passes up to each pixel's last blended entry.

    Vectorised per tile: alpha [P, G] for the tile's P pixels and G list entries, exclusive cumprod for the
    transmittance, the T <= 1e-4 stop rule as a cumulative mask.  Differentiable (masks are constants, the
    0.999 clamp passes gradient only when not clamped — same as the CUDA backward).
    """
    C, N = opacities.shape
    D = colors.shape[-1]
    dt = means2d.dtype
    tile_h, tile_w = isect_offsets.shape[1], isect_offsets.shape[2]
    n_isects = flatten_ids.numel()
    m2 = means2d.reshape(C * N, 2)
    cn = conics.reshape(C * N, 3)
    cl = colors.reshape(C * N, D)
    op = opacities.reshape(C * N)
    offs = isect_offsets.reshape(-1).tolist() + [n_isects]
    out = torch.zeros(C, height, width, D, dtype=dt)
    alphas = torch.zeros(C, height, width, 1, dtype=dt)
    last_ids = torch.zeros(C, height, width, dtype=torch.int32)
    rows_c, rows_a = [[None] * (tile_h * tile_w) for _ in range(C)], None
    out_tiles = {}
    for c in range(C):
        for ty in range(tile_h):
            for tx in range(tile_w):
                lin = (c * tile_h + ty) * tile_w + tx
                s, e = offs[lin], offs[lin + 1]
                y0, x0 = ty * tile_size, tx * tile_size
                y1, x1 = min(y0 + tile_size, height), min(x0 + tile_size, width)
                ph, pw = y1 - y0, x1 - x0
                if e <= s:
                    col = torch.zeros(ph, pw, D, dtype=dt)
                    a = torch.zeros(ph, pw, 1, dtype=dt)
                    T_fin = torch.ones(ph, pw, 1, dtype=dt)
                else:
                    g = flatten_ids[s:e].long()
                    py, px = torch.meshgrid(torch.arange(y0, y1, dtype=dt) + 0.5, torch.arange(x0, x1, dtype=dt) + 0.5,
                                            indexing="ij")
                    px, py = px.reshape(-1, 1), py.reshape(-1, 1)
                    dx = m2[g, 0][None] - px
                    dy = m2[g, 1][None] - py
                    sigma = 0.5 * (cn[g, 0][None] * dx * dx + cn[g, 2][None] * dy * dy) + cn[g, 1][None] * dx * dy
                    alpha = torch.clamp(op[g][None] * torch.exp(-sigma), max=ALPHA_MAX)
                    valid = (sigma >= 0) & (alpha >= ALPHA_MIN)
                    a_eff = torch.where(valid, alpha, torch.zeros_like(alpha))
                    one_m = 1.0 - a_eff
                    T_incl = torch.cumprod(one_m, dim=1)
                    T_excl = torch.cat([torch.ones_like(T_incl[:, :1]), T_incl[:, :-1]], dim=1)
                    stop = valid & (T_incl.detach() <= T_MIN)
                    dead = torch.cumsum(stop.to(torch.int32), dim=1) > 0
                    contrib = valid & ~dead
                    w = torch.where(contrib, a_eff * T_excl, torch.zeros_like(a_eff))
                    col = (w @ cl[g]).reshape(ph, pw, D)
                    log_keep = torch.where(contrib, one_m, torch.ones_like(one_m))
                    T_fin = torch.prod(log_keep, dim=1).reshape(ph, pw, 1)
                    a = 1.0 - T_fin
                    idx = torch.arange(s, e, dtype=torch.int32)[None].expand_as(contrib)
                    li = torch.where(contrib, idx, torch.zeros_like(idx)).max(dim=1).values
                    last_ids[c, y0:y1, x0:x1] = li.reshape(ph, pw)
                    if stats is not None:
                        stats["blended"] = stats.get("blended", 0) + int(contrib.sum())
                        stats["visited"] = stats.get("visited", 0) + int(torch.clamp(li.long() - s + 1, min=0).sum())
                if backgrounds is not None:
                    col = col + T_fin * backgrounds[c][None, None, :]
                out_tiles[(c, ty, tx)] = (col, a)
    # stitch (keeps autograd): rows of tiles -> image
    imgs_c, imgs_a = [], []
    for c in range(C):
        rows_col, rows_al = [], []
        for ty in range(tile_h):
            rows_col.append(torch.cat([out_tiles[(c, ty, tx)][0] for tx in range(tile_w)], dim=1))
            rows_al.append(torch.cat([out_tiles[(c, ty, tx)][1] for tx in range(tile_w)], dim=1))
        imgs_c.append(torch.cat(rows_col, dim=0))
        imgs_a.append(torch.cat(rows_al, dim=0))
    out = torch.stack(imgs_c, dim=0)
    alphas = torch.stack(imgs_a, dim=0)
    if return_last_ids:
        return out, alphas, last_ids
    return out, alphas


# --------------------------------------------------------------------------------------------
# A.1 rasterization()   (unpacked path, as called from dn_model.py:570-591)
# --------------------------------------------------------------------------------------------
def rasterization(means, quats, scales, opacities, colors, viewmats, Ks, width, height, near_plane=0.01,
                  far_plane=1e10, radius_clip=0.0, eps2d=0.3, sh_degree=None, packed=False, tile_size=16,
                  backgrounds=None, render_mode="RGB", sparse_grad=False, absgrad=False, rasterize_mode="classic",
                  channel_chunk=32) -> Tuple[Tensor, Tensor, Dict]:
    assert not packed and not sparse_grad
    N, C = means.shape[0], viewmats.shape[0]
    radii, means2d, depths, conics, comps = fully_fused_projection(
        means, quats, scales, viewmats, Ks, width, height, eps2d, near_plane, far_plane, radius_clip,
        calc_compensations=(rasterize_mode == "antialiased"))
    opac = opacities[None].repeat(C, 1)
    if comps is not None:
        opac = opac * comps
    tile_w, tile_h = math.ceil(width / tile_size), math.ceil(height / tile_size)
    with torch.no_grad():
        tiles_per_gauss, isect_ids, flatten_ids = isect_tiles(means2d, radii, depths, tile_size, tile_w, tile_h)
        isect_offsets = isect_offset_encode(isect_ids, C, tile_w, tile_h)
    if sh_degree is not None and render_mode not in ("D", "ED"):
        camtoworlds = torch.linalg.inv(viewmats)
        dirs = means[None, :, :] - camtoworlds[:, None, :3, 3]
        cols = spherical_harmonics(sh_degree, dirs, colors[None].expand(C, *colors.shape), masks=radii > 0)
        cols = torch.clamp_min(cols + 0.5, 0.0)
    elif render_mode in ("D", "ED"):
        cols = None
    else:
        cols = colors if colors.dim() == 3 else colors[None].expand(C, N, colors.shape[-1])
    if render_mode in ("RGB+D", "RGB+ED"):
        cols = torch.cat([cols, depths[..., None]], dim=-1)
    elif render_mode in ("D", "ED"):
        cols = depths[..., None]
    render_colors, render_alphas = rasterize_to_pixels(means2d, conics, cols, opac, width, height, tile_size,
                                                       isect_offsets, flatten_ids, backgrounds=backgrounds)
    if render_mode in ("ED", "RGB+ED"):
        render_colors = torch.cat(
            [render_colors[..., :-1], render_colors[..., -1:] / render_alphas.clamp(min=1e-10)], dim=-1)
    meta = {
        "camera_ids": None, "gaussian_ids": None, "radii": radii, "means2d": means2d, "depths": depths,
        "conics": conics, "opacities": opac, "tile_width": tile_w, "tile_height": tile_h,
        "tiles_per_gauss": tiles_per_gauss, "isect_ids": isect_ids, "flatten_ids": flatten_ids,
        "isect_offsets": isect_offsets, "width": width, "height": height, "tile_size": tile_size, "n_cameras": C,
    }
    return render_colors, render_alphas, meta


# --------------------------------------------------------------------------------------------
# A.6 legacy rasterize_gaussians  (normals pass, dn_model.py:644-653)
# --------------------------------------------------------------------------------------------
def rasterize_gaussians(xys, depths, radii, conics, num_tiles_hit, colors, opacity, img_height, img_width,
                        block_width, background=None, return_alpha=False):
    N, D = colors.shape
    if background is None:
        background = torch.ones(D, dtype=colors.dtype)
    tile_w, tile_h = math.ceil(img_width / block_width), math.ceil(img_height / block_width)
    with torch.no_grad():
        _, ids, flat = isect_tiles(xys[None].detach(), radii.reshape(1, N), depths.reshape(1, N).detach(),
                                   block_width, tile_w, tile_h, legacy_bbox=True)
        offs = isect_offset_encode(ids, 1, tile_w, tile_h)
    if ids.numel() < 1:
        out = torch.ones(img_height, img_width, D, dtype=colors.dtype) * background
        alpha = torch.zeros(img_height, img_width, dtype=colors.dtype)
    else:
        o, a = rasterize_to_pixels(xys[None], conics[None], colors[None], opacity.reshape(1, N), img_width,
                                   img_height, block_width, offs, flat, backgrounds=background[None])
        out, alpha = o[0], a[0, ..., 0]
    return (out, alpha) if return_alpha else out
