"""GPU: intersection lists pruned by the exact reach test (csrc/isect_reach.cu, DNSplatterStepConfig.prune_lists;
first run on a B200 in round 2: profiles/r02a_pytest_prune.log).

What has to hold: the pruned sorted list is the full sorted list with entries removed (same order, same keys); no
removed entry reaches a pixel; images and gradients of the render are unchanged; the captured step agrees with the
unpruned one."""
import math
import os

import numpy as np
import pytest
import torch

from fusionsense_b200.synthetic import make_scene
from tests.parity import assert_close

pytestmark = [pytest.mark.gpu]
DEV = "cuda"


def _projected(n=20000, W=320, H=240, cfg_id=71, kind="bunny"):
    from fusionsense_b200 import ops

    sc = make_scene(n, W, H, n_views=3, cfg_id=cfg_id, kind=kind, fx=300.0).to(DEV)
    coeffs = torch.cat((sc.features_dc[:, None, :], sc.features_rest), dim=1)
    q = sc.quats / sc.quats.norm(dim=-1, keepdim=True)
    radii, m2, dep, con, _, _, tiles = ops.project_sh_fwd(
        sc.means, q, torch.exp(sc.scales), sc.viewmats[:1].contiguous(), sc.Ks[:1].contiguous(), W, H, 0.3, 0.01, 1e10,
        0.0, 16, 3, coeffs.contiguous(), None, 4, 3, False)
    opac = torch.sigmoid(sc.opacities[:, 0])[None].contiguous()
    return sc, radii, m2, dep, con, opac, tiles


@pytest.mark.parametrize("legacy", [False, True])
@pytest.mark.parametrize("kind", ["bunny", "random"])
def test_pruned_list_is_the_full_list_minus_unreachable_pairs(legacy, kind):
    from fusionsense_b200 import ops

    W, H = 320, 240
    sc, radii, m2, dep, con, opac, tiles = _projected(W=W, H=H, kind=kind)
    tw, th = math.ceil(W / 16), math.ceil(H / 16)
    _, ids_f, flat_f, offs_f = ops.isect_tiles(m2, radii, dep, 16, tw, th, legacy_bbox=legacy)
    _, ids_p, flat_p, offs_p = ops.isect_tiles(m2, radii, dep, 16, tw, th, legacy_bbox=legacy, reach=(con, opac))
    ids_f, flat_f, ids_p, flat_p = (t.cpu().numpy() for t in (ids_f, flat_f, ids_p, flat_p))
    assert 0 < len(ids_p) < len(ids_f)
    # subsequence check: walk the full list, consume the pruned one in order
    keep = np.zeros(len(ids_f), dtype=bool)
    j = 0
    for i in range(len(ids_f)):
        if j < len(ids_p) and ids_f[i] == ids_p[j] and flat_f[i] == flat_p[j]:
            keep[i] = True
            j += 1
    assert j == len(ids_p), "pruned list is not an order-preserving subsequence of the full list"
    # every dropped pair fails the alpha test on every pixel centre of its tile (float64 check)
    m2n, conn, opn = m2[0].cpu().numpy().astype(np.float64), con[0].cpu().numpy().astype(np.float64), opac[0].cpu().numpy()
    tile = (ids_f >> 32) & ((1 << ops.tile_bits_for(tw * th)) - 1)
    drop = np.nonzero(~keep)[0]
    yy, xx = np.meshgrid(np.arange(16) + 0.5, np.arange(16) + 0.5, indexing="ij")
    for i in drop[:: max(1, len(drop) // 4000)]:
        g, t = flat_f[i], tile[i]
        ty, tx = divmod(int(t), tw)
        dx, dy = m2n[g, 0] - (xx + tx * 16), m2n[g, 1] - (yy + ty * 16)
        sigma = 0.5 * (conn[g, 0] * dx * dx + conn[g, 2] * dy * dy) + conn[g, 1] * dx * dy
        alpha = np.minimum(0.999, opn[g] * np.exp(-sigma))
        assert not ((sigma >= 0) & (alpha >= 1 / 255)).any(), (i, g, t)
    # ranges agree with the pruned keys
    offs = offs_p.reshape(-1).cpu().numpy().tolist() + [len(ids_p)]
    tile_p = (ids_p >> 32) & ((1 << ops.tile_bits_for(tw * th)) - 1)
    for t in range(tw * th):
        assert (tile_p[offs[t]:offs[t + 1]] == t).all()


@pytest.mark.parametrize("sh_degree", [3, 0])
def test_render_and_gradients_do_not_change(sh_degree):
    from fusionsense_b200.gsplat import rasterization_from_params

    sc = make_scene(9000, 320, 240, n_views=3, cfg_id=61, kind="bunny", fx=300.0).to(DEV)
    g = torch.Generator().manual_seed(5)
    cot = (torch.rand(1, 240, 320, 4, generator=g) * 2 - 1).to(DEV)
    cot_a = (torch.rand(1, 240, 320, 1, generator=g) * 2 - 1).to(DEV)
    res = []
    for prune in (True, False):
        P = {k: getattr(sc, k).clone().requires_grad_(True) for k in
             ("means", "quats", "scales", "opacities", "features_dc", "features_rest")}
        render, alpha, info = rasterization_from_params(
            P["means"], P["quats"], P["scales"], torch.sigmoid(P["opacities"]).squeeze(-1), P["features_dc"],
            P["features_rest"], sc.viewmats[:1], sc.Ks[:1], 320, 240, sh_degree=sh_degree, render_mode="RGB+ED",
            absgrad=True, prune_lists=prune)
        info["means2d"].retain_grad()
        ((render * cot).sum() + (alpha * cot_a).sum()).backward()
        res.append((render.detach(), alpha.detach(), info, {k: v.grad for k, v in P.items()}))
    (ra, aa, ia, ga), (rb, ab, ib, gb) = res
    assert ia["flatten_ids"].numel() < ib["flatten_ids"].numel()
    assert torch.equal(ia["tiles_per_gauss"], ib["tiles_per_gauss"])
    assert_close(ra, rb, "prune.render", tol=1e-6, outlier_frac=1e-5)
    assert_close(aa, ab, "prune.alpha", tol=1e-6, outlier_frac=1e-5)
    assert_close(ia["means2d"].absgrad, ib["means2d"].absgrad, "prune.absgrad", tol=1e-5, outlier_frac=1e-4)
    for k in ga:
        assert_close(ga[k], gb[k], f"prune.grad.{k}", tol=1e-5, outlier_frac=1e-4)


def test_step_eager_and_captured_agree_with_unpruned():
    from fusionsense_b200.dn_step import DNSplatterStep, DNSplatterStepConfig
    from fusionsense_b200.graph_step import GraphedDNSplatterStep

    sc = make_scene(20000, 320, 240, n_views=3, cfg_id=51, kind="bunny", fx=300.0)
    on = DNSplatterStep(sc, DNSplatterStepConfig(prune_lists=True, fused_passes=False), device=DEV, step=3000)
    off = DNSplatterStep(sc, DNSplatterStepConfig(prune_lists=False, fused_passes=False), device=DEV, step=3000)
    batch = off.render_targets(1)
    outs = {}
    for name, m in (("on", on), ("off", off)):
        out = m.get_outputs(0)
        loss = m.get_loss_dict(out, batch)["main_loss"]
        loss.backward()
        outs[name] = (out, loss)
    for k in ("rgb", "depth", "normal", "accumulation"):
        assert_close(outs["on"][0][k], outs["off"][0][k], f"prune.step.out.{k}", tol=1e-6, outlier_frac=1e-5)
    assert float(outs["on"][1]) == pytest.approx(float(outs["off"][1]), rel=1e-6)
    for k in on.gauss_params:
        assert_close(on.gauss_params[k].grad, off.gauss_params[k].grad, f"prune.step.grad.{k}", tol=1e-5,
                     outlier_frac=1e-4)
    # captured step (static-capacity mode: device-side counts, shared lists for the normals pass)
    targets = {v: off.render_targets(v) for v in range(3)}
    runs = {}
    for name, prune in (("on", True), ("off", False)):
        m = DNSplatterStep(sc, DNSplatterStepConfig(prune_lists=prune, fused_passes=False), device=DEV, step=3000)
        r = GraphedDNSplatterStep(m, targets)
        for i in range(4):
            r.train_iteration(i % 3)
        info = r.poll()
        assert info["overflowed_steps"] == 0
        runs[name] = (m, info)
    assert runs["on"][1]["n_isects"] < runs["off"][1]["n_isects"]
    assert runs["on"][1]["n_isects"] == runs["on"][1]["n_isects_normals"]
    assert runs["on"][1]["loss"] == pytest.approx(runs["off"][1]["loss"], rel=1e-5)
    for k in runs["on"][0].gauss_params:
        assert_close(runs["on"][0].gauss_params[k].data, runs["off"][0].gauss_params[k].data,
                     f"prune.graph.param.{k}", tol=1e-5, outlier_frac=1e-4)


@pytest.mark.parametrize("begin,end,n", [(32, 45, 300000), (32, 43, 5000), (0, 32, 70000), (32, 51, 40000), (5, 6, 1000)])
def test_radix_sort_keys_window_is_a_stable_sort_on_those_bits(begin, end, n):
    """fsb_radix_sort_keys against torch's stable sort of the extracted bit window; the low words ride along."""
    from fusionsense_b200 import ops

    g = torch.Generator().manual_seed(begin * 64 + end)
    keys = torch.randint(0, 2**62, (n,), generator=g, dtype=torch.int64)
    # few distinct high digits and long runs of equal ones, as tile ids have
    keys[: n // 2] &= ~(((1 << (end - begin)) - 1) << begin) | (0x15 << begin)
    keys = keys.to(DEV)
    window = (keys >> begin) & ((1 << (end - begin)) - 1)
    order = torch.sort(window, stable=True).indices
    want = keys[order]
    got, low = ops.radix_sort_keys(keys.clone(), begin, end)
    assert torch.equal(got, want)
    assert torch.equal(low.long() & 0xFFFFFFFF, want & 0xFFFFFFFF)
    # static-capacity form: only the first *n_dev keys are sorted
    m = n // 3
    n_dev = torch.tensor([m], dtype=torch.int64, device=DEV)
    got2, low2 = ops.radix_sort_keys(keys.clone(), begin, end, n_dev=n_dev)
    order2 = torch.sort(window[:m], stable=True).indices
    assert torch.equal(got2[:m], keys[:m][order2])
    assert torch.equal(low2[:m].long() & 0xFFFFFFFF, keys[:m][order2] & 0xFFFFFFFF)


@pytest.mark.parametrize("legacy", [False, 2])
@pytest.mark.parametrize("C", [1, 3])
@pytest.mark.parametrize("static", [False, True])
def test_two_level_binning_gives_the_order_of_the_64_bit_key_sort(legacy, C, static, monkeypatch):
    """Depth sort of the Gaussians + emission in that order + stable sort on the (camera, tile) bits == one stable sort
    of the (camera | tile | depth bits) keys: same flatten ids (flags included), same keys, same tile ranges.  Half of
    the Gaussians are exact duplicates of the other half so equal depths inside a tile are common and the tie rule
    (emission order = Gaussian index) is exercised."""
    from fusionsense_b200 import ops

    W, H = 320, 240
    sc = make_scene(6000, W, H, n_views=3, cfg_id=72, kind="bunny", fx=300.0).to(DEV)
    dup = lambda t: torch.cat([t, t], dim=0).contiguous()
    coeffs = dup(torch.cat((sc.features_dc[:, None, :], sc.features_rest), dim=1))
    q = dup(sc.quats / sc.quats.norm(dim=-1, keepdim=True))
    radii, m2, dep, con, _, _, tiles = ops.project_sh_fwd(
        dup(sc.means), q, dup(torch.exp(sc.scales)), sc.viewmats[:C].contiguous(), sc.Ks[:C].contiguous(), W, H, 0.3,
        0.01, 1e10, 0.0, 16, 3, coeffs, None, 4, 3, False)
    opac = dup(torch.sigmoid(sc.opacities[:, 0]))[None].expand(C, -1).contiguous()
    tw, th = math.ceil(W / 16), math.ceil(H / 16)

    def run(two_level):
        monkeypatch.setattr(ops, "TWO_LEVEL_BINNING", two_level)
        if not static:
            _, ids, flat, offs = ops.isect_tiles(m2, radii, dep, 16, tw, th, legacy_bbox=legacy, reach=(con, opac))
            return ids, flat, offs
        overflow = torch.zeros(1, dtype=torch.int32, device=DEV)
        with ops.static_capacity(400000, overflow):
            _, ids, flat, offs = ops.isect_tiles(m2, radii, dep, 16, tw, th, legacy_bbox=legacy, reach=(con, opac))
        n = int(flat.n_dev.item())
        assert int(overflow.item()) == 0 and 0 < n < 400000
        return ids[:n], flat[:n], offs

    ids1, flat1, offs1 = run(False)
    ids2, flat2, offs2 = run(True)
    assert ids1.numel() > 20000
    assert torch.equal(flat1, flat2)
    assert torch.equal(ids1, ids2)
    assert torch.equal(offs1, offs2)
    # ties exist: the same (tile, depth) key on neighbouring entries
    assert int((ids1[1:] == ids1[:-1]).sum()) > 1000
