"""GPU parity at BASELINE.json's FULL sizes (cfg3 512^3 hull, cfg4 1M Gaussians @ 1920x1080, cfg5 3M @ 3840x2160),
where the CPU oracle cannot run the whole problem in seconds.  Two kinds of checks:

  * size-independent properties: sortedness and stability of the intersection lists, offsets consistent with the
    keys, sum(tiles_per_gauss) == n_isects, compositing linear in the colours, the adjoint identity
    <v_out, R(c)> == <R^T(v_out), c> of the raster backward, slab-sharded hull == unsharded hull;
  * oracle-anchored windows: a tile-aligned 64x64 window of the full image only sees the Gaussians in its tiles'
    lists, so the same Gaussians translated by the window origin form a small scene the oracle renders and
    differentiates in milliseconds; the full-size forward inside the window and the full-size backward of a
    cotangent supported on the window must match it (tests/parity.py tolerances).  For the hull, whole z-planes of
    the 512^3 grid are voted by the numpy oracle and must match bit for bit.
"""
import math

import numpy as np
import pytest
import torch

from fusionsense_b200.synthetic import make_scene
from oracle import gsplat_ref as ref
from oracle import visual_hull_ref as vh_ref
from tests.golden_io import load_visual_hull_golden, write_visual_hull_capture
from tests.parity import assert_close

pytestmark = pytest.mark.gpu
DEV = "cuda"
TS = 16


def _project_and_bin(n, W, H, cfg_id, scale_mult):
    from fusionsense_b200 import ops

    sc = make_scene(n, W, H, n_views=2, cfg_id=cfg_id, kind="random").to(DEV)
    coeffs = torch.cat([sc.features_dc[:, None, :], sc.features_rest], dim=1).contiguous()
    scales = (torch.exp(sc.scales) * scale_mult).contiguous()
    radii, m2, dep, con, _comp, cols, tiles = ops.project_sh_fwd(
        sc.means, sc.quats, scales, sc.viewmats[:1].contiguous(), sc.Ks[:1].contiguous(), W, H, 0.3, 0.01, 1e10, 0.0,
        TS, 3, coeffs, None, 4, 3, False)
    tw, th = math.ceil(W / TS), math.ceil(H / TS)
    _, ids, flat, offs = ops.isect_tiles(m2, radii, dep, TS, tw, th, tiles_per_gauss=tiles)
    opac = torch.sigmoid(sc.opacities[:, 0])[None].contiguous()
    return ops, dict(radii=radii, m2=m2, dep=dep, con=con, cols=cols.contiguous(), tiles=tiles, ids=ids, flat=flat,
                     offs=offs, opac=opac, tw=tw, th=th, W=W, H=H, n=n)


def _check_lists(s):
    ids, flat, offs, tiles = s["ids"], s["flat"], s["offs"], s["tiles"]
    I = ids.numel()
    assert I == int(tiles.sum()) and I > 4 * s["n"] // 10
    # sorted on the unsigned key; ties (same tile, same depth bits) keep emission order = ascending Gaussian index
    assert bool((ids[1:] >= ids[:-1]).all())  # keys are < 2^63, signed compare is the unsigned one
    ties = ids[1:] == ids[:-1]
    assert bool((flat[1:][ties] > flat[:-1][ties]).all())
    # key = tile << 32 | depth bits of that Gaussian
    dep_bits = s["dep"].reshape(-1).view(torch.int32).long()
    assert torch.equal(ids & 0xFFFFFFFF, dep_bits[flat.long()])
    tile_of = (ids >> 32)
    n_tiles = s["tw"] * s["th"]
    assert int(tile_of.max()) < n_tiles
    # offsets[t] = first sorted position whose tile id is >= t
    expect = torch.searchsorted(tile_of.contiguous(), torch.arange(n_tiles, device=DEV))
    assert torch.equal(offs.reshape(-1).long(), expect)
    # every Gaussian appears exactly tiles_per_gauss times
    assert torch.equal(torch.bincount(flat.long(), minlength=s["n"]), tiles.reshape(-1).long())


def _window_vs_oracle(ops, s, ox, oy, win=64, seed=0, tag=""):
    """Forward inside, and backward of a cotangent supported on, the tile-aligned window [oy, oy+win) x [ox, ox+win)."""
    W, H = s["W"], s["H"]
    hh, ww = min(win, H - oy), min(win, W - ox)  # a window over the ragged last tile row / column is clipped
    g = torch.Generator().manual_seed(seed)
    out, alpha, last, ws = ops.raster_fwd(s["m2"], s["con"], s["cols"], s["opac"], None, None, W, H, TS, s["offs"],
                                          s["flat"])
    v_out = torch.zeros_like(out)
    v_alpha = torch.zeros_like(alpha)
    v_out[0, oy:oy + hh, ox:ox + ww] = torch.randn(win, win, 4, generator=g)[:hh, :ww].to(DEV)
    v_alpha[0, oy:oy + hh, ox:ox + ww] = torch.randn(win, win, 1, generator=g)[:hh, :ww].to(DEV)
    v_m2, v_abs, v_con, v_col, v_op = ops.raster_bwd(s["m2"], s["con"], s["cols"], s["opac"], None, None, W, H, TS,
                                                     s["offs"], s["flat"], False, ws, out, alpha, last, v_out, v_alpha,
                                                     True)
    # the Gaussians in the window's tile lists
    offs = torch.cat([s["offs"].reshape(-1).long(), torch.tensor([s["flat"].numel()], device=DEV)])
    sel = []
    for ty in range(oy // TS, min((oy + win) // TS, s["th"])):
        for tx in range(ox // TS, min((ox + win) // TS, s["tw"])):
            t = ty * s["tw"] + tx
            sel.append(s["flat"][offs[t]:offs[t + 1]])
    gids = torch.unique(torch.cat(sel).long())  # ascending: ties in the sub-scene break the same way
    assert gids.numel() > 50, "window must not be empty"
    # nothing outside the window's lists may have received a gradient
    touched = torch.zeros(s["n"], dtype=torch.bool, device=DEV)
    touched[gids] = True
    assert float(v_col[0][~touched].abs().max()) == 0.0 and float(v_op[0][~touched].abs().max()) == 0.0

    shift = torch.tensor([float(ox), float(oy)])
    m2s = (s["m2"][0, gids].cpu() - shift)[None].requires_grad_(True)
    cons = s["con"][0, gids].cpu()[None].requires_grad_(True)
    cols = s["cols"][0, gids].cpu()[None].requires_grad_(True)
    opas = s["opac"][0, gids].cpu()[None].requires_grad_(True)
    radii = s["radii"][0, gids].cpu()[None]
    deps = s["dep"][0, gids].cpu()[None]
    tw = win // TS
    _, ids_s, flat_s = ref.isect_tiles(m2s.detach(), radii, deps, TS, tw, tw)
    offs_s = ref.isect_offset_encode(ids_s, 1, tw, tw)
    o_ref, a_ref = ref.rasterize_to_pixels(m2s, cons, cols, opas, win, win, TS, offs_s, flat_s)
    o_ref, a_ref = o_ref[:, :hh, :ww], a_ref[:, :hh, :ww]
    assert_close(out[0, oy:oy + hh, ox:ox + ww].cpu(), o_ref[0].detach(), f"full.fwd.colors{tag}", tol=1e-4)
    assert_close(alpha[0, oy:oy + hh, ox:ox + ww].cpu(), a_ref[0].detach(), f"full.fwd.alpha{tag}", tol=1e-4)
    loss = (o_ref * v_out[:, oy:oy + hh, ox:ox + ww].cpu()).sum() + (a_ref * v_alpha[:, oy:oy + hh, ox:ox + ww].cpu()).sum()
    loss.backward()
    assert_close(v_col[0, gids].cpu(), cols.grad[0], f"full.bwd.v_colors{tag}", tol=1e-4, outlier_frac=2e-3)
    assert_close(v_op[0, gids].cpu(), opas.grad[0], f"full.bwd.v_opacities{tag}", tol=1e-4, outlier_frac=2e-3)
    assert_close(v_con[0, gids].cpu(), cons.grad[0], f"full.bwd.v_conics{tag}", tol=1e-4, outlier_frac=2e-3)
    assert_close(v_m2[0, gids].cpu(), m2s.grad[0], f"full.bwd.v_means2d{tag}", tol=1e-4, outlier_frac=2e-3)
    assert bool((v_abs[0, gids] >= v_m2[0, gids].abs() * (1 - 1e-4) - 1e-12).all())  # sum |g| >= |sum g|


def _linearity_and_adjoint(ops, s, seed=1):
    W, H, n = s["W"], s["H"], s["n"]
    g = torch.Generator().manual_seed(seed)
    c1 = torch.rand(1, n, 3, generator=g).to(DEV)
    c2 = torch.rand(1, n, 3, generator=g).to(DEV)
    args = (s["opac"], None, None, W, H, TS, s["offs"], s["flat"])
    o1, a1, l1, _ = ops.raster_fwd(s["m2"], s["con"], c1, *args)
    o2, a2, l2, _ = ops.raster_fwd(s["m2"], s["con"], c2, *args)
    o12, a12, l12, ws = ops.raster_fwd(s["m2"], s["con"], (c1 + 2 * c2).contiguous(), *args)
    # geometry does not depend on the colours: identical alpha and last ids, bit for bit
    assert torch.equal(a1, a2) and torch.equal(a1, a12) and torch.equal(l1, l2) and torch.equal(l1, l12)
    assert float(a1.min()) >= 0.0 and float(a1.max()) < 1.0
    assert_close(o12.cpu(), (o1 + 2 * o2).cpu(), "full.linearity", tol=1e-5, outlier_frac=1e-4)
    # adjoint identity of the colour gradient (the raster is linear in the colours)
    v_out = torch.randn(o12.shape, generator=g).to(DEV)
    _, _, _, v_col, _ = ops.raster_bwd(s["m2"], s["con"], (c1 + 2 * c2).contiguous(), s["opac"], None, None, W, H, TS,
                                       s["offs"], s["flat"], False, ws, o12, a12, l12, v_out, torch.zeros_like(a12),
                                       False)
    lhs = float((v_out.double() * o1.double()).sum())
    rhs = float((v_col.double() * c1.double()).sum())
    scale = float((v_out.double().abs() * o1.double()).sum())
    assert abs(lhs - rhs) <= 1e-5 * scale, (lhs, rhs, scale)


def test_cfg4_one_million_gaussians_1080p():
    ops, s = _project_and_bin(1_000_000, 1920, 1080, cfg_id=4, scale_mult=1.0)
    _check_lists(s)
    _linearity_and_adjoint(ops, s)
    _window_vs_oracle(ops, s, ox=960, oy=512, tag="[cfg4,centre]")
    # ragged bottom rows: 1080 = 67.5 tiles, the window covers tile rows 64..67 of which the last has 8 pixel rows
    _window_vs_oracle(ops, s, ox=928, oy=1024, seed=3, tag="[cfg4,bottom]")


def test_cfg5_three_million_gaussians_4k():
    ops, s = _project_and_bin(3_000_000, 3840, 2160, cfg_id=5, scale_mult=1.0)
    assert s["tw"] * s["th"] == 32400  # 15 tile bits
    _check_lists(s)
    _linearity_and_adjoint(ops, s)
    _window_vs_oracle(ops, s, ox=1920, oy=1024, tag="[cfg5,centre]")


def test_cfg3_visual_hull_512_cubed(tmp_path):
    """512^3 voxels, 9 masks 640x480: the numpy oracle votes whole z-planes (first, middle, last of every slab);
    slab-sharded carving (world 4) equals the unsharded run bit for bit."""
    from fusionsense_b200 import visual_hull as vh

    g = load_visual_hull_golden()
    path = write_visual_hull_capture(tmp_path, g)
    mats, centre, names = vh.read_hull_cameras(path)
    masks = vh.read_masks(path, names)
    n = 512
    xs, ys, zs = vh.hull_grid(centre, half_extent=0.5, n_per_axis=n)
    whole = vh.HullCarver(mats, masks, xs, ys, zs)
    maxv = whole.vote()
    iso = vh.iso_value(maxv, 5)
    pts, idx = whole.extract(iso, want_indices=True)
    assert whole.votes.numel() == n ** 3 and pts.shape[0] > 1000
    plane = n * n
    for iz in (0, 1, 127, 128, 255, 256, 383, 384, 511):
        v_ref = vh_ref.project_votes(mats, masks, xs, ys, zs[iz:iz + 1])
        assert np.array_equal(whole.votes[iz * plane:(iz + 1) * plane].cpu().numpy(), v_ref), iz
    # occupied set == threshold of the votes, in voxel order, with the grid's coordinates
    occ = torch.nonzero(whole.votes > iso).reshape(-1)
    assert torch.equal(idx, occ)
    iz, rem = occ // plane, occ % plane
    ix, iy = rem // n, rem % n
    expect = torch.stack([torch.from_numpy(xs).to(DEV)[ix], torch.from_numpy(ys).to(DEV)[iy],
                          torch.from_numpy(zs).to(DEV)[iz]], dim=1)
    assert torch.equal(pts, expect)
    # slabs
    votes_whole = whole.votes
    parts, maxes = [], []
    for r in range(4):
        c = vh.HullCarver(mats, masks, xs, ys, zs, rank=r, world_size=4)
        maxes.append(c.vote())
        z0, z1 = vh.slab_bounds(n, r, 4)
        assert torch.equal(c.votes, votes_whole[z0 * plane:z1 * plane])
        parts.append(c.extract(iso))
        del c
    assert max(maxes) == maxv
    assert torch.equal(torch.cat(parts), pts)
