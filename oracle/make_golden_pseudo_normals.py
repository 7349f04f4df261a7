"""Generate tests/golden/pseudo_normals.npz from the UNMODIFIED reference functions
/root/reference/dn_splatter/utils/normal_utils.py::normal_from_depth_image / pcd_to_normal (with
/root/reference/dn_splatter/utils/camera_utils.py), imported in this container.

Run from the repo root:  python -m oracle.make_golden_pseudo_normals
The real dn_splatter/__init__.py imports nerfstudio (absent here), so the two utility modules are loaded under a
fake `dn_splatter` package object; neither touches nerfstudio.
"""
import importlib.util
import sys
import types
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parent.parent
REF = Path("/root/reference/dn_splatter")


def load_reference_normal_utils():
    pkg = types.ModuleType("dn_splatter")
    pkg.__path__ = []
    utils = types.ModuleType("dn_splatter.utils")
    utils.__path__ = []
    sys.modules["dn_splatter"], sys.modules["dn_splatter.utils"] = pkg, utils
    mods = {}
    for name in ("camera_utils", "normal_utils"):
        spec = importlib.util.spec_from_file_location(f"dn_splatter.utils.{name}", REF / "utils" / f"{name}.py")
        mod = importlib.util.module_from_spec(spec)
        sys.modules[f"dn_splatter.utils.{name}"] = mod
        spec.loader.exec_module(mod)
        mods[name] = mod
    return mods["normal_utils"]


def main():
    nu = load_reference_normal_utils()
    g = torch.Generator().manual_seed(20241013)
    W, H = 53, 37
    yy, xx = torch.meshgrid(torch.arange(H, dtype=torch.float32), torch.arange(W, dtype=torch.float32), indexing="ij")
    # a tilted plane plus a bump and a little noise; a few zero-depth holes as a RealSense frame has
    depth = 0.4 + 0.002 * xx - 0.003 * yy + 0.05 * torch.exp(-((xx - 25) ** 2 + (yy - 18) ** 2) / 60.0)
    depth = depth + 0.001 * torch.rand(H, W, generator=g)
    depth[torch.rand(H, W, generator=g) < 0.03] = 0.0
    depth = depth[..., None].contiguous()
    fx, fy, cx, cy = 61.5, 60.25, 26.3, 18.9
    eye = torch.eye(4)
    n_eye = nu.normal_from_depth_image(depth, fx, fy, cx, cy, (W, H), eye, torch.device("cpu"))
    ang = 0.7
    c2w = torch.tensor([[np.cos(ang), 0.0, np.sin(ang), 0.3], [0.0, 1.0, 0.0, -0.2],
                        [-np.sin(ang), 0.0, np.cos(ang), 1.1], [0.0, 0.0, 0.0, 1.0]], dtype=torch.float32)
    n_pose = nu.normal_from_depth_image(depth, fx, fy, cx, cy, (W, H), c2w, torch.device("cpu"))
    xyz = torch.randn(H, W, 3, generator=g)
    n_pcd = nu.pcd_to_normal(xyz)
    # the way dn_model.py:775-795 turns it into a [0,1] ground-truth normal image
    gt = (1 + n_eye @ torch.diag(torch.tensor([1.0, -1.0, -1.0]))) / 2
    out = ROOT / "tests" / "golden" / "pseudo_normals.npz"
    np.savez_compressed(out, depth=depth.numpy(), intr=np.array([fx, fy, cx, cy], dtype=np.float64),
                        c2w=c2w.numpy(), n_eye=n_eye.numpy(), n_pose=n_pose.numpy(), xyz=xyz.numpy(),
                        n_pcd=n_pcd.numpy(), gt_normal=gt.numpy())
    print(f"wrote {out} ({out.stat().st_size} bytes)")


if __name__ == "__main__":
    main()
