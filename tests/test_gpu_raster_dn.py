"""GPU: the two compositing passes of a DN-Splatter iteration in one walk (fsb_raster_dn_fwd / _bwd, csrc/raster.cu)
against (i) the same kernels run as two separate passes on their own lists and (ii) the CPU oracle.

Colour set A = rasterization() semantics on the gsplat 1.0 lists (dn_model.py:570-591), colour set B = the legacy
rasterize_gaussians semantics on the 0.1.x lists with a white background (dn_model.py:644-653).  The union list holds
the 0.1.x-only (Gaussian, tile) pairs with FSB_LEGACY_FLAG; the second test builds a scene in which such pairs exist
and visibly contribute, so the flagged-tile path is exercised."""
import math

import pytest
import torch

from fusionsense_b200.synthetic import make_scene
from oracle import gsplat_ref as ref
from tests.parity import assert_close

pytestmark = pytest.mark.gpu
DEV = "cuda"


def _two_pass(ops, m2, con, col4, nrm, opac, radii, dep, W, H, v_a, v_b, v_al, prune):
    """Reference structure: two plain compositing passes, each on its own sorted list."""
    tw, th = math.ceil(W / 16), math.ceil(H / 16)
    reach = (con, opac) if prune else None
    _, _, flat_a, offs_a = ops.isect_tiles(m2, radii, dep, 16, tw, th, reach=reach)
    _, _, flat_b, offs_b = ops.isect_tiles(m2, radii, dep, 16, tw, th, legacy_bbox=True, reach=reach)
    m2a = m2.clone().requires_grad_(True)
    cona, col4a, opa = con.clone().requires_grad_(True), col4.clone().requires_grad_(True), opac.clone().requires_grad_(True)
    nrma = nrm.clone().requires_grad_(True)
    out_a, al = ops.RasterizeToPixels.apply(m2a, cona, col4a, opa, None, None, W, H, 16, offs_a, flat_a, True, True)
    bg = torch.ones(1, 3, device=DEV)
    out_b, _ = ops.RasterizeToPixels.apply(m2a.detach(), cona, nrma, opa, bg, None, W, H, 16, offs_b, flat_b, False, False)
    loss = (out_a * v_a).sum() + (out_b * v_b).sum() + (al * v_al).sum()
    loss.backward()
    return (out_a.detach(), out_b.detach(), al.detach(),
            dict(means2d=m2a.grad, absgrad=m2a.absgrad, conics=cona.grad, colors_a=col4a.grad, colors_b=nrma.grad,
                 opacities=opa.grad), (flat_a, offs_a, flat_b, offs_b))


def _one_pass(ops, m2, con, col4, nrm, opac, radii, dep, W, H, v_a, v_b, v_al):
    tw, th = math.ceil(W / 16), math.ceil(H / 16)
    _, _, flat_u, offs_u = ops.isect_tiles(m2, radii, dep, 16, tw, th, legacy_bbox=2, reach=(con, opac))
    m2a = m2.clone().requires_grad_(True)
    cona, col4a, opa = con.clone().requires_grad_(True), col4.clone().requires_grad_(True), opac.clone().requires_grad_(True)
    nrma = nrm.clone().requires_grad_(True)
    bg = torch.ones(1, 3, device=DEV)
    out_a, out_b, al = ops.RasterizeDN.apply(m2a, cona, col4a, nrma, opa, None, bg, W, H, 16, offs_u, flat_u, True, 3)
    loss = (out_a * v_a).sum() + (out_b * v_b).sum() + (al * v_al).sum()
    loss.backward()
    return (out_a.detach(), out_b.detach(), al.detach(),
            dict(means2d=m2a.grad, absgrad=m2a.absgrad, conics=cona.grad, colors_a=col4a.grad, colors_b=nrma.grad,
                 opacities=opa.grad), (flat_u, offs_u))


def _compare(tag, one, two):
    for name, a, b in (("rgbd", one[0], two[0]), ("normals", one[1], two[1]), ("alpha", one[2], two[2])):
        assert_close(a, b, f"raster_dn.{tag}.{name}", tol=1e-5, outlier_frac=1e-4)
    for k in one[3]:
        assert_close(one[3][k], two[3][k], f"raster_dn.{tag}.grad.{k}", tol=1e-4, outlier_frac=2e-3)


@pytest.mark.parametrize("n,W,H,kind,scale_mult", [(20000, 320, 240, "bunny", 1.0), (30000, 640, 480, "random", 3.0),
                                                   (40000, 128, 96, "random", 8.0)])
def test_one_walk_equals_two_passes(n, W, H, kind, scale_mult):
    """Projected scenes (light tiles, and with scale_mult 8 on 128x96 several-thousand-entry segment-parallel tiles)."""
    from fusionsense_b200 import ops

    sc = make_scene(n, W, H, n_views=2, cfg_id=81, kind=kind, fx=300.0 if kind == "bunny" else None).to(DEV)
    coeffs = torch.cat((sc.features_dc[:, None, :], sc.features_rest), dim=1).contiguous()
    q = sc.quats / sc.quats.norm(dim=-1, keepdim=True)
    radii, m2, dep, con, _, col4, tiles = ops.project_sh_fwd(
        sc.means, q, (torch.exp(sc.scales) * scale_mult).contiguous(), sc.viewmats[:1].contiguous(),
        sc.Ks[:1].contiguous(), W, H, 0.3, 0.01, 1e10, 0.0, 16, 3, coeffs, None, 4, 3, False)
    opac = torch.sigmoid(sc.opacities[:, 0])[None].contiguous()
    g = torch.Generator().manual_seed(5)
    nrm = torch.nn.functional.normalize(torch.randn(1, n, 3, generator=g), dim=-1).to(DEV)
    v_a = torch.randn(1, H, W, 4, generator=g).to(DEV)
    v_b = torch.randn(1, H, W, 3, generator=g).to(DEV)
    v_al = torch.randn(1, H, W, 1, generator=g).to(DEV)
    two = _two_pass(ops, m2, con, col4, nrm, opac, radii, dep, W, H, v_a, v_b, v_al, prune=True)
    one = _one_pass(ops, m2, con, col4, nrm, opac, radii, dep, W, H, v_a, v_b, v_al)
    _compare(f"{kind}{n}", one, two)
    # and the pruned two-pass structure equals the unpruned one (the reference's lists)
    two_full = _two_pass(ops, m2, con, col4, nrm, opac, radii, dep, W, H, v_a, v_b, v_al, prune=False)
    _compare(f"{kind}{n}.vs_full_lists", one, two_full)


def _flagged_scene(W=128, H=96, n=600, seed=3):
    """2-D Gaussians placed so that means2d.x + radius is an exact multiple of 16 for a third of them: the 0.1.x box
    rule ((int)(x + r) / 16 + 1) then yields one tile column more than the 1.0 rule (ceil), and with a large radius
    and opacity ~1 that column is reached with alpha > 1/255."""
    g = torch.Generator().manual_seed(seed)
    radii = torch.randint(6, 40, (n,), generator=g, dtype=torch.int32)
    m2 = torch.stack([torch.rand(n, generator=g) * W, torch.rand(n, generator=g) * H], dim=-1)
    k = torch.randint(1, W // 16, (n,), generator=g).float()
    snap = torch.arange(n) % 3 == 0
    m2[snap, 0] = (16.0 * k - radii.float())[snap]
    sig = radii.float() / 3.0
    # isotropic-ish conics with a little correlation; sigma = radius / 3 along x
    a = 1.0 / sig**2
    c = a * (0.6 + 0.8 * torch.rand(n, generator=g))
    b = 0.3 * torch.sqrt(a * c) * (torch.rand(n, generator=g) * 2 - 1)
    con = torch.stack([a, b, c], dim=-1)
    opac = torch.where(snap, torch.full((n,), 0.995), 0.05 + 0.9 * torch.rand(n, generator=g))
    dep = 0.5 + torch.rand(n, generator=g) * 4
    col4 = torch.cat([torch.rand(n, 3, generator=g), dep[:, None]], dim=-1)
    nrm = torch.nn.functional.normalize(torch.randn(n, 3, generator=g), dim=-1)
    return [t[None].contiguous() for t in (radii, m2, dep, con, opac, col4, nrm)]


def test_flagged_entries_reach_set_b_only():
    from fusionsense_b200 import ops

    W, H = 128, 96
    radii, m2, dep, con, opac, col4, nrm = [t.to(DEV) for t in _flagged_scene(W, H)]
    g = torch.Generator().manual_seed(11)
    v_a = torch.randn(1, H, W, 4, generator=g).to(DEV)
    v_b = torch.randn(1, H, W, 3, generator=g).to(DEV)
    v_al = torch.randn(1, H, W, 1, generator=g).to(DEV)
    two = _two_pass(ops, m2, con, col4, nrm, opac, radii, dep, W, H, v_a, v_b, v_al, prune=True)
    one = _one_pass(ops, m2, con, col4, nrm, opac, radii, dep, W, H, v_a, v_b, v_al)
    flat_u = one[4][0].cpu()
    n_flag = int((flat_u < 0).sum())
    assert n_flag > 0, "the scene must produce 0.1.x-only list entries"
    flat_a, _, flat_b, _ = two[4]
    assert flat_u.numel() == flat_b.numel() and flat_u.numel() - n_flag == flat_a.numel()
    assert torch.equal(flat_u & 0x7FFFFFFF, flat_b.cpu())  # the union list IS the legacy list, flags aside
    _compare("flagged", one, two)
    # sensitivity: the flagged entries change the normals image (set B on the 1.0 lists would differ)
    tw, th = math.ceil(W / 16), math.ceil(H / 16)
    bg = torch.ones(1, 3, device=DEV)
    out_b_on_a_lists, _ = ops.RasterizeToPixels.apply(m2, con, nrm, opac, bg, None, W, H, 16, two[4][1], two[4][0],
                                                      False, False)
    assert (out_b_on_a_lists - two[1]).abs().max() > 1e-4
    # both colour sets against the CPU oracle on its own two lists
    m2c, conc, opc, radc, depc = m2.cpu(), con.cpu(), opac.cpu(), radii.cpu(), dep.cpu()
    _, ids_a, fl_a = ref.isect_tiles(m2c, radc, depc, 16, tw, th)
    offs_a = ref.isect_offset_encode(ids_a, 1, tw, th)
    _, ids_b, fl_b = ref.isect_tiles(m2c, radc, depc, 16, tw, th, legacy_bbox=True)
    offs_b = ref.isect_offset_encode(ids_b, 1, tw, th)
    oa, al = ref.rasterize_to_pixels(m2c, conc, col4.cpu(), opc, W, H, 16, offs_a, fl_a)
    oa = torch.cat([oa[..., :3], oa[..., 3:] / al.clamp(min=1e-10)], dim=-1)
    ob, _ = ref.rasterize_to_pixels(m2c, conc, nrm.cpu(), opc, W, H, 16, offs_b, fl_b, backgrounds=torch.ones(1, 3))
    assert_close(one[0], oa, "raster_dn.flagged.rgbd_vs_oracle", tol=1e-4, outlier_frac=1e-3)
    assert_close(one[1], ob, "raster_dn.flagged.normals_vs_oracle", tol=1e-4, outlier_frac=1e-3)
    assert_close(one[2], al, "raster_dn.flagged.alpha_vs_oracle", tol=1e-4, outlier_frac=1e-3)


def test_step_fused_passes_matches_separate_passes():
    """DNSplatterStep: fused_passes (one walk) against the two-pass structure, outputs, loss and parameter gradients."""
    from fusionsense_b200.dn_step import DNSplatterStep, DNSplatterStepConfig

    scene = make_scene(20000, 320, 240, n_views=3, cfg_id=83, kind="bunny", fx=300.0)
    res = {}
    for name, fp in (("one", True), ("two", False)):
        m = DNSplatterStep(scene, DNSplatterStepConfig(fused_passes=fp, prune_lists=False), device=DEV, step=3000)
        batch = m.render_targets(1)
        out = m.get_outputs(0)
        loss = m.get_loss_dict(out, batch)["main_loss"]
        loss.backward()
        res[name] = (out, float(loss), {k: p.grad.clone() for k, p in m.gauss_params.items()})
    for k in ("rgb", "depth", "normal", "accumulation"):
        assert_close(res["one"][0][k], res["two"][0][k], f"step.fused_passes.{k}", tol=1e-5, outlier_frac=1e-4)
    assert res["one"][1] == pytest.approx(res["two"][1], rel=1e-5)
    for k in res["one"][2]:
        assert_close(res["one"][2][k], res["two"][2][k], f"step.fused_passes.grad.{k}", tol=1e-4, outlier_frac=2e-3)
