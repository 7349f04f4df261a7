"""The plain-torch loss restatement (oracle/dn_losses_ref.py, CPU) and the fused CUDA regulariser (GPU) against
outputs of the unmodified reference dn_splatter/losses.py (tests/golden/dn_losses.npz,
generator: oracle/make_golden_losses.py)."""
import numpy as np
import pytest
import torch

from tests.golden_io import GOLDEN


def _load(device="cpu"):
    z = np.load(GOLDEN / "dn_losses.npz")
    t = {k: torch.from_numpy(z[k]).to(device) for k in ("depth", "sensor", "rgb", "pred_normal", "gt_normal", "pred_rgb")}
    return z, t


def test_mirrored_torch_classes_match_reference_outputs():
    from fusionsense_b200.losses import DepthLossType
    from oracle.dn_losses_ref import DepthLoss, TVLoss

    z, t = _load()
    gt_img = t["rgb"].clamp(min=10 / 255.0)
    valid = t["sensor"] > 0.1
    assert DepthLoss(DepthLossType.EdgeAwareLogL1)(t["depth"], t["sensor"], gt_img, valid).item() == pytest.approx(
        float(z["ea_logl1"]), rel=1e-6)
    assert DepthLoss(DepthLossType.TV)(t["depth"]).item() == pytest.approx(float(z["tv_depth"]), rel=1e-6)
    assert TVLoss()(t["pred_normal"]).item() == pytest.approx(float(z["tv_normal"]), rel=1e-6)
    assert DepthLoss(DepthLossType.LogL1)(t["depth"], t["sensor"]).item() == pytest.approx(float(z["logl1"]), rel=1e-6)
    assert DepthLoss(DepthLossType.EdgeAwareTV)(t["depth"], t["rgb"]).item() == pytest.approx(float(z["ea_tv"]), rel=1e-6)


def test_ssim_restatement_properties():
    from oracle.dn_losses_ref import SSIM

    g = torch.Generator().manual_seed(1)
    a = torch.rand(1, 3, 48, 64, generator=g)
    s = SSIM()
    assert s(a, a).item() == pytest.approx(1.0, abs=1e-6)
    b = (a + 0.1 * torch.randn(a.shape, generator=g)).clamp(0, 1)
    v = s(a, b).item()
    assert 0.0 < v < 1.0
    assert s(b, a).item() == pytest.approx(v, rel=1e-6)  # symmetric


@pytest.mark.gpu
def test_fused_regulariser_matches_reference_golden():
    from fusionsense_b200.losses import dn_regularizer_loss

    z, t = _load("cuda")
    ins = {k: t[k].clone().requires_grad_(True) for k in ("depth", "pred_normal", "pred_rgb")}
    loss = dn_regularizer_loss(ins["depth"], t["sensor"], t["rgb"], ins["pred_normal"], t["gt_normal"], ins["pred_rgb"],
                               t["rgb"], depth_tolerance=0.1, sensor_depth_lambda=0.2, smooth_loss_lambda=0.1,
                               normal_l1_lambda=0.4, normal_tv_lambda=0.4, rgb_l1_lambda=0.8)
    assert loss.item() == pytest.approx(float(z["total"]), rel=2e-6)
    (3.0 * loss).backward()  # a non-unit upstream gradient, read from device memory by the kernel
    for name, key in (("depth", "v_depth"), ("pred_normal", "v_pred_normal"), ("pred_rgb", "v_pred_rgb")):
        ref = 3.0 * z[key]
        got = ins[name].grad.cpu().numpy()
        assert np.abs(got - ref).max() <= 1e-5 * np.abs(ref).max() + 1e-12, name


@pytest.mark.gpu
def test_fused_regulariser_single_terms_and_full_frame():
    """Each term alone against the mirrored torch classes on a 640x480 frame (the bench size)."""
    from fusionsense_b200.losses import DepthLossType, dn_regularizer_loss
    from oracle.dn_losses_ref import DepthLoss, TVLoss

    g = torch.Generator().manual_seed(3)
    H, W = 480, 640
    depth = (0.2 + torch.rand(H, W, 1, generator=g)).cuda()
    sensor = (0.2 + torch.rand(H, W, 1, generator=g))
    sensor[torch.rand(H, W, 1, generator=g) < 0.2] = 0
    sensor = sensor.cuda()
    rgb = torch.rand(H, W, 3, generator=g).cuda()
    pn = torch.rand(H, W, 3, generator=g).cuda()
    gn = torch.rand(H, W, 3, generator=g).cuda()
    kw = dict(sensor_depth_lambda=0.0, smooth_loss_lambda=0.0, normal_l1_lambda=0.0, normal_tv_lambda=0.0)
    cases = {
        "sensor_depth_lambda": lambda d, n: DepthLoss(DepthLossType.EdgeAwareLogL1)(d, sensor, rgb.clamp(min=10 / 255.0), sensor > 0.1),
        "smooth_loss_lambda": lambda d, n: DepthLoss(DepthLossType.TV)(d),
        "normal_l1_lambda": lambda d, n: torch.abs(gn - n).mean(),
        "normal_tv_lambda": lambda d, n: TVLoss()(n),
    }
    for key, fn in cases.items():
        d1, n1 = depth.clone().requires_grad_(True), pn.clone().requires_grad_(True)
        d2, n2 = depth.clone().requires_grad_(True), pn.clone().requires_grad_(True)
        fused = dn_regularizer_loss(d1, sensor, rgb, n1, gn, **{**kw, key: 1.0})
        ref = fn(d2, n2)
        assert fused.item() == pytest.approx(ref.item(), rel=5e-6), key
        fused.backward()
        ref.backward()
        for a, b in ((d1, d2), (n1, n2)):
            ga = a.grad if a.grad is not None else torch.zeros_like(a)
            gb = b.grad if b.grad is not None else torch.zeros_like(b)
            assert (ga - gb).abs().max() <= 1e-5 * gb.abs().max() + 1e-12, key
