"""Generate tests/golden/dn_losses.npz by importing the UNMODIFIED reference module
/root/reference/dn_splatter/losses.py (and dn_splatter/utils/normal_utils.py) in this container.

Run from the repo root:  python -m oracle.make_golden_losses
torchmetrics and nerfstudio are not installed; losses.py only imports names from them at module top
(losses.py:11-16) without using them in the classes exercised here, so they are stubbed in sys.modules.
"""
import importlib.util
import sys
import types
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parent.parent
REF = Path("/root/reference")


def load_reference_losses():
    tm = types.ModuleType("torchmetrics")
    tmi = types.ModuleType("torchmetrics.image")
    tmi.MultiScaleStructuralSimilarityIndexMeasure = object
    tmi.StructuralSimilarityIndexMeasure = object
    tm.image = tmi
    ns = types.ModuleType("nerfstudio")
    fc = types.ModuleType("nerfstudio.field_components")
    fh = types.ModuleType("nerfstudio.field_components.field_heads")
    fh.FieldHeadNames = object
    for name, mod in (("torchmetrics", tm), ("torchmetrics.image", tmi), ("nerfstudio", ns),
                      ("nerfstudio.field_components", fc), ("nerfstudio.field_components.field_heads", fh)):
        sys.modules.setdefault(name, mod)
    spec = importlib.util.spec_from_file_location("ref_dn_losses", REF / "dn_splatter" / "losses.py")
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def main():
    L = load_reference_losses()
    g = torch.Generator().manual_seed(20241011)
    H, W = 37, 53
    depth = (0.3 + torch.rand(H, W, 1, generator=g)).requires_grad_(True)
    sensor = 0.3 + torch.rand(H, W, 1, generator=g)
    sensor[torch.rand(H, W, 1, generator=g) < 0.3] = 0.0  # invalid sensor pixels (<= depth_tolerance)
    rgb = torch.rand(H, W, 3, generator=g)
    pred_n = torch.rand(H, W, 3, generator=g).requires_grad_(True)
    gt_n = torch.rand(H, W, 3, generator=g)
    pred_rgb = torch.rand(H, W, 3, generator=g).requires_grad_(True)
    valid = sensor > 0.1
    gt_img = rgb.clamp(min=10 / 255.0)

    ea = L.DepthLoss(L.DepthLossType.EdgeAwareLogL1)(depth, sensor.float(), gt_img, valid)
    tv_d = L.DepthLoss(L.DepthLossType.TV)(depth)
    nl1 = torch.abs(gt_n - pred_n).mean()
    tv_n = L.TVLoss()(pred_n)
    l1_rgb = torch.abs(rgb - pred_rgb).mean()
    logl1 = L.DepthLoss(L.DepthLossType.LogL1)(depth, sensor)
    eatv = L.DepthLoss(L.DepthLossType.EdgeAwareTV)(depth, rgb)
    # FusionSense weighting (configs/config.py:10-11, dn_model.py:68,74): 0.2, 0.1, normal_lambda 0.4; base L1 0.8
    total = 0.2 * ea + 0.1 * tv_d + 0.4 * (nl1 + tv_n) + 0.8 * l1_rgb
    total.backward()
    out = ROOT / "tests" / "golden" / "dn_losses.npz"
    np.savez_compressed(
        out, depth=depth.detach().numpy(), sensor=sensor.numpy(), rgb=rgb.numpy(), pred_normal=pred_n.detach().numpy(),
        gt_normal=gt_n.numpy(), pred_rgb=pred_rgb.detach().numpy(), ea_logl1=ea.item(), tv_depth=tv_d.item(),
        normal_l1=nl1.item(), tv_normal=tv_n.item(), l1_rgb=l1_rgb.item(), logl1=logl1.item(), ea_tv=eatv.item(),
        total=total.item(), v_depth=depth.grad.numpy(), v_pred_normal=pred_n.grad.numpy(),
        v_pred_rgb=pred_rgb.grad.numpy())
    print(f"wrote {out} ({out.stat().st_size} bytes): total={total.item():.8f}")


if __name__ == "__main__":
    main()
