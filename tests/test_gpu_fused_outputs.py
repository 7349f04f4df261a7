"""GPU: the fused get_outputs path (rasterization_from_params, compose_rgbd, normal_map, flatness_loss) against the
torch expressions of the reference they replace (dn_model.py:566-574, :602-613, :655-656, :817-819), values and
gradients, and the whole step with the switch on against the same step with it off."""
import pytest
import torch

from fusionsense_b200.synthetic import make_scene
from tests.parity import assert_close

pytestmark = pytest.mark.gpu


def _rand(*shape, seed=0, lo=0.0, hi=1.0):
    g = torch.Generator().manual_seed(seed)
    return (torch.rand(*shape, generator=g) * (hi - lo) + lo).cuda()


@pytest.mark.parametrize("H,W", [(48, 64), (37, 53), (1, 5)])
def test_compose_rgbd_matches_torch(H, W):
    from fusionsense_b200.compose import compose_rgbd

    render = _rand(1, H, W, 4, seed=1, lo=-0.2, hi=1.3)
    alpha = _rand(1, H, W, 1, seed=2)
    alpha[0, :: 3, :: 2] = 0.0  # empty pixels take the detached maximum
    render[..., 3] = render[..., 3].abs() * 5 * (alpha[..., 0] > 0)
    bg = torch.tensor([1.0, 0.5, 0.25]).cuda()
    v_rgb, v_depth = _rand(H, W, 3, seed=3, lo=-1, hi=1), _rand(H, W, 1, seed=4, lo=-1, hi=1)

    def torch_form(render, alpha):
        rgb = torch.clamp(render[:, ..., :3] + (1 - alpha) * bg, 0.0, 1.0)
        d = render[:, ..., 3:4]
        d = torch.where(alpha > 0, d, d.detach().max()).squeeze(0)
        return rgb.squeeze(0), d

    outs = []
    for fn in (lambda r, a: compose_rgbd(r, a, bg), torch_form):
        r, a = render.clone().requires_grad_(True), alpha.clone().requires_grad_(True)
        rgb, depth = fn(r, a)
        ((rgb * v_rgb).sum() + (depth * v_depth).sum()).backward()
        outs.append((rgb.detach(), depth.detach(), r.grad, a.grad))
    for name, x, y in zip(("rgb", "depth", "v_render", "v_alpha"), *outs):
        assert x.shape == y.shape, (name, x.shape, y.shape)
        assert torch.equal(x, y) or float((x - y).abs().max()) <= 1e-6 * float(y.abs().max() + 1e-12), name
    # only one of the two cotangents present
    r, a = render.clone().requires_grad_(True), alpha.clone().requires_grad_(True)
    rgb, depth = compose_rgbd(r, a, bg)
    (depth * v_depth).sum().backward()
    assert float(r.grad[..., :3].abs().max()) == 0.0 and float(a.grad.abs().max()) == 0.0


def test_compose_rgbd_all_empty_and_negative_depths():
    from fusionsense_b200.compose import compose_rgbd

    bg = torch.ones(3).cuda()
    render = torch.zeros(1, 8, 8, 4).cuda()
    render[..., 3] = -_rand(1, 8, 8, seed=5, lo=0.5, hi=2.0)  # all negative: the maximum is the least negative one
    alpha = torch.zeros(1, 8, 8, 1).cuda()
    rgb, depth = compose_rgbd(render, alpha, bg)
    assert torch.equal(rgb, torch.ones(8, 8, 3).cuda())
    assert torch.equal(depth, render[0, ..., 3:4].max().expand(8, 8, 1))


def test_normal_map_matches_torch():
    from fusionsense_b200.compose import normal_map

    n = _rand(33, 47, 3, seed=6, lo=-1, hi=1)
    v = _rand(33, 47, 3, seed=7, lo=-1, hi=1)
    a = n.clone().requires_grad_(True)
    out_a = normal_map(a)
    (out_a * v).sum().backward()
    b = n.clone().requires_grad_(True)
    t = b / b.norm(dim=-1, keepdim=True)
    out_b = (t + 1) / 2
    (out_b * v).sum().backward()
    assert_close(out_a, out_b, "normal_map.out", tol=1e-6, outlier_frac=0)
    assert_close(a.grad, b.grad, "normal_map.grad", tol=1e-5, outlier_frac=1e-4)


@pytest.mark.parametrize("N", [1, 257, 100_003])
def test_flatness_loss_matches_torch(N):
    from fusionsense_b200.compose import flatness_loss

    s = _rand(N, 3, seed=8, lo=-7, hi=-2)
    a = s.clone().requires_grad_(True)
    la = flatness_loss(a)
    (la * 0.37).backward()
    b = s.clone().requires_grad_(True)
    lb = torch.min(torch.exp(b), dim=1, keepdim=True)[0].mean()
    (lb * 0.37).backward()
    assert float(la) == pytest.approx(float(lb), rel=2e-6)
    assert_close(a.grad, b.grad, "flatness.grad", tol=1e-6, outlier_frac=0)


def test_combine_losses_matches_torch_bitwise():
    from fusionsense_b200.compose import combine_losses

    vals = [torch.tensor(v, device="cuda", requires_grad=True) for v in (0.7312345, 0.0412345, 0.0034567)]
    ref = [v.detach().clone().requires_grad_(True) for v in vals]
    out = combine_losses(vals[0], vals[1], vals[2], 0.2, 0.4)
    (out * 0.5).backward()
    exp = 0.2 * (1 - ref[0]) + ref[1] + 0.4 * (0 + ref[2])
    (exp * 0.5).backward()
    assert float(out) == float(exp)
    for a, b in zip(vals, ref):
        assert float(a.grad) == pytest.approx(float(b.grad), rel=1e-7)
    # absent terms
    o2 = combine_losses(vals[0].detach(), None, None, 0.2, 0.4)
    assert float(o2) == float(0.2 * (1 - ref[0].detach()))


@pytest.mark.parametrize("sh_degree,C", [(3, 1), (1, 1), (0, 1), (2, 2)])
def test_rasterization_from_params_matches_rasterization(sh_degree, C):
    from fusionsense_b200.gsplat import rasterization, rasterization_from_params

    sc = make_scene(9000, 320, 240, n_views=3, cfg_id=61, kind="bunny", fx=300.0).to("cuda")
    vm, Ks = sc.viewmats[:C], sc.Ks[:C]
    cot = _rand(C, 240, 320, 4, seed=9, lo=-1, hi=1)
    cot_a = _rand(C, 240, 320, 1, seed=10, lo=-1, hi=1)
    res = []
    for fused in (True, False):
        P = {k: getattr(sc, k).clone().requires_grad_(True) for k in
             ("means", "quats", "scales", "opacities", "features_dc", "features_rest")}
        opac = torch.sigmoid(P["opacities"]).squeeze(-1)
        if fused:
            render, alpha, info = rasterization_from_params(
                P["means"], P["quats"], P["scales"], opac, P["features_dc"], P["features_rest"], vm, Ks, 320, 240,
                sh_degree=sh_degree, render_mode="RGB+ED", absgrad=True)
        else:
            q = P["quats"]
            render, alpha, info = rasterization(
                means=P["means"], quats=q / q.norm(dim=-1, keepdim=True), scales=torch.exp(P["scales"]),
                opacities=opac, colors=torch.cat((P["features_dc"][:, None, :], P["features_rest"]), dim=1),
                viewmats=vm, Ks=Ks, width=320, height=240, packed=False, render_mode="RGB+ED", sh_degree=sh_degree,
                absgrad=True)
        info["means2d"].retain_grad()
        ((render * cot).sum() + (alpha * cot_a).sum()).backward()
        res.append((render.detach(), alpha.detach(), info, {k: v.grad for k, v in P.items()}))
    (ra, aa, ia, ga), (rb, ab, ib, gb) = res
    assert torch.equal(ia["radii"], ib["radii"])
    assert torch.equal(ia["tiles_per_gauss"], ib["tiles_per_gauss"])
    assert torch.equal(ia["flatten_ids"], ib["flatten_ids"]) and torch.equal(ia["isect_offsets"], ib["isect_offsets"])
    assert_close(ra, rb, "from_params.render", tol=1e-5, outlier_frac=1e-4)
    assert_close(aa, ab, "from_params.alpha", tol=1e-5, outlier_frac=1e-4)
    assert_close(ia["means2d"].absgrad, ib["means2d"].absgrad, "from_params.absgrad", tol=1e-4, outlier_frac=2e-3)
    for k in ga:
        assert ga[k] is not None and ga[k].shape == gb[k].shape, k
        assert_close(ga[k], gb[k], f"from_params.grad.{k}", tol=1e-4, outlier_frac=2e-3)
    # bands above the active degree receive exactly zero
    nb = (sh_degree + 1) ** 2
    if nb < 16:
        assert float(ga["features_rest"][:, nb - 1:, :].abs().max()) == 0.0


@pytest.mark.parametrize("step", [3000, 1500, 0])
def test_step_with_fused_outputs_matches_step_without(step):
    from fusionsense_b200.dn_step import DNSplatterStep, DNSplatterStepConfig

    sc = make_scene(20000, 320, 240, n_views=3, cfg_id=51, kind="bunny", fx=300.0)
    on = DNSplatterStep(sc, DNSplatterStepConfig(fused_outputs=True), device="cuda", step=step)
    off = DNSplatterStep(sc, DNSplatterStepConfig(fused_outputs=False), device="cuda", step=step)
    batch = off.render_targets(1)
    outs = {}
    for name, m in (("on", on), ("off", off)):
        out = m.get_outputs(0)
        loss = m.get_loss_dict(out, batch)["main_loss"]
        loss.backward()
        outs[name] = (out, loss)
    for k in ("rgb", "depth", "normal", "accumulation"):
        assert outs["on"][0][k].shape == outs["off"][0][k].shape, k
        assert_close(outs["on"][0][k], outs["off"][0][k], f"fused_outputs.out.{k}", tol=1e-5, outlier_frac=1e-4)
    assert float(outs["on"][1]) == pytest.approx(float(outs["off"][1]), rel=1e-5)
    for k in on.gauss_params:
        assert_close(on.gauss_params[k].grad, off.gauss_params[k].grad, f"fused_outputs.grad.{k}", tol=1e-4,
                     outlier_frac=2e-3)
    assert_close(on.xys.absgrad, off.xys.absgrad, "fused_outputs.absgrad", tol=1e-4, outlier_frac=2e-3)


def test_batched_render_views_match_single_camera_outputs():
    """eval_render.render_views (C cameras per rasterization call, both colour sets in one walk) against
    get_outputs in eval mode, view by view."""
    from fusionsense_b200.dn_step import DNSplatterStep, DNSplatterStepConfig
    from fusionsense_b200.eval_render import render_views
    from fusionsense_b200.synthetic import make_scene
    from tests.parity import assert_close

    sc = make_scene(15000, 256, 192, n_views=5, cfg_id=67, kind="bunny", fx=240.0)
    m = DNSplatterStep(sc, DNSplatterStepConfig(), device="cuda", step=3000)
    m.training = False
    singles = []
    with torch.no_grad():
        for v in range(5):
            singles.append({k: t.clone() for k, t in m.get_outputs(v).items()})
    batched = list(render_views(m, range(5), chunk=3))
    assert [b["view"] for b in batched] == list(range(5))
    for v in range(5):
        for k in ("rgb", "depth", "normal", "accumulation"):
            assert_close(batched[v][k], singles[v][k], f"render_views.{v}.{k}", tol=1e-6, outlier_frac=1e-5)
