"""GPU: fused per-step image metrics (csrc/metrics.cu, fusionsense_b200/metrics.py) against the formulas of the
reference's dn_splatter/metrics.py (restated in torch here; the reference's own classes are exercised through the
harness when baseline/_ref is installed)."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def _inputs(H=240, W=320, seed=0):
    g = torch.Generator().manual_seed(seed)
    gt = torch.rand(H, W, 3, generator=g).cuda()
    pred = (gt + 0.05 * torch.randn(H, W, 3, generator=g).cuda()).clamp(0, 1)
    gd = (0.05 + 2 * torch.rand(H, W, 1, generator=g)).cuda()
    gd[torch.rand(H, W, 1, generator=g).cuda() < 0.3] = 0.0  # invalid sensor pixels
    pd = (gd + 0.1 * torch.randn(H, W, 1, generator=g).cuda()).abs() + 0.01
    return pred, gt, pd, gd


def _depth_metrics_torch(pred, gt, tol=0.1):
    pred, gt = pred.double(), gt.double()
    mask = gt > tol
    thresh = torch.max(gt[mask] / pred[mask], pred[mask] / gt[mask])
    a = [(thresh < 1.25**k).double().mean() for k in (1, 2, 3)]
    rmse = torch.sqrt(((gt[mask] - pred[mask]) ** 2).mean())
    rmse_log = torch.sqrt((torch.log(gt[mask]) - torch.log(pred[mask])) ** 2).nanmean()
    abs_rel = (torch.abs(gt - pred)[mask] / gt[mask]).mean()
    sq_rel = ((gt - pred)[mask] ** 2 / gt[mask]).mean()
    return [abs_rel, sq_rel, rmse, rmse_log] + a


def test_image_metrics_match_reference_formulas():
    from fusionsense_b200.metrics import NAMES, image_metrics

    pred, gt, pd, gd = _inputs()
    for _ in range(2):  # the workspace is left zeroed: a second launch gives the same numbers
        m = image_metrics(pred, gt, pd.squeeze(-1), gd.squeeze(-1), 0.1).cpu().double()
    mse = ((pred.double() - gt.double()) ** 2).mean()
    assert float(m[0]) == pytest.approx(float(mse), rel=1e-5)
    assert float(m[1]) == pytest.approx(float(10 * torch.log10(1.0 / mse)), rel=1e-5)
    ref = _depth_metrics_torch(pd, gd)
    for i, r in enumerate(ref):
        assert float(m[2 + i]) == pytest.approx(float(r), rel=2e-5), NAMES[2 + i]
    assert int(m[9]) == int((gd > 0.1).sum())


def test_metric_classes_keep_the_reference_signatures():
    from fusionsense_b200.metrics import DepthMetrics, RGBMetrics
    from oracle.dn_losses_ref import SSIM

    pred, gt, pd, gd = _inputs(seed=1)
    calls = []

    def fake_lpips(p, g):
        calls.append(1)
        return (p - g).abs().mean()

    rgb = RGBMetrics(lpips=fake_lpips, lpips_every=3)
    for i in range(4):
        psnr, ssim, lp = rgb(pred.permute(2, 0, 1)[None], gt.permute(2, 0, 1)[None])
    assert len(calls) == 2  # calls 0 and 3
    ref_ssim = SSIM().cuda().double()(pred.double().permute(2, 0, 1)[None], gt.double().permute(2, 0, 1)[None])
    assert float(ssim) == pytest.approx(float(ref_ssim), rel=1e-5)
    assert float(lp) == pytest.approx(float((pred - gt).abs().mean()), rel=1e-6)
    d = DepthMetrics(tolerance=0.1)(pd.permute(2, 0, 1), gd.permute(2, 0, 1))
    ref = _depth_metrics_torch(pd, gd)
    for x, r in zip(d, ref):
        assert float(x) == pytest.approx(float(r), rel=2e-5)


def test_step_metrics_in_the_captured_step():
    from fusionsense_b200.dn_step import DNSplatterStep, DNSplatterStepConfig
    from fusionsense_b200.graph_step import GraphedDNSplatterStep
    from fusionsense_b200.synthetic import make_scene

    sc = make_scene(8000, 256, 192, n_views=3, cfg_id=61, kind="bunny", fx=240.0)
    m = DNSplatterStep(sc, DNSplatterStepConfig(step_metrics=True), device="cuda", step=3000)
    targets = {v: m.render_targets(v) for v in range(3)}
    runner = GraphedDNSplatterStep(m, targets)
    seen = []
    for v in (0, 1, 2, 0):
        runner.train_iteration(v)
        runner.poll()
        seen.append({k: float(t) for k, t in m.last_metrics.items()})
    assert runner.captures == 1
    assert all(10.0 < s["rgb_psnr"] < 80.0 and 0.0 < s["rgb_ssim"] <= 1.0 for s in seen)
    assert seen[0]["rgb_psnr"] != seen[1]["rgb_psnr"]  # rewritten by every replay (a different view)
    assert seen[3]["gaussian_count"] == 8000
