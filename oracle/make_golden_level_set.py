"""Generate tests/golden/level_set.npz by running the UNMODIFIED reference function
`/root/reference/dn_splatter/dn_model.py::DNSplatterModel.compute_level_surface_points` in this container, on the CPU:
the unbound function is called with a stand-in `self` that carries what the function reads (Gaussian parameters,
`normals`, `config.knn_to_track`, `device`, and a `get_outputs` that returns a fixed synthetic depth / colour image —
the render itself needs the GPU and is not what this fixture pins).  `knn_sk` is the reference's own (sklearn);
`random.sample` is replaced by "the first k in order" for the duration of the call so the rows stay in pixel order.

dn_model.py is imported against the stub nerfstudio / torchmetrics packages of tests/stubs and this repository's gsplat
shim (`quat_to_rotmat`, plain torch).  Run from the repo root:  python -m oracle.make_golden_level_set
"""
import random
import sys
import types
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parent.parent
REF = Path("/root/reference")


def scene(seed=5, n=4000, H=48, W=64):
    g = torch.Generator().manual_seed(seed)
    # Gaussians on a sphere of radius 0.3 (plus a loose cloud behind it), camera on +z looking at the origin (OpenGL c2w)
    u = torch.nn.functional.normalize(torch.randn(n, 3, generator=g), dim=-1)
    means = 0.3 * u + 0.004 * torch.randn(n, 3, generator=g)
    means[: n // 10] = torch.randn(n // 10, 3, generator=g) * 0.5 + torch.tensor([0.0, 0.0, -1.0])
    log_scales = torch.log(0.02 * torch.exp(0.4 * torch.randn(n, 3, generator=g)))
    quats = torch.randn(n, 4, generator=g)
    opac = 1.0 + 1.5 * torch.randn(n, 1, generator=g)
    normals = torch.nn.functional.normalize(torch.randn(n, 3, generator=g), dim=-1)
    cam = torch.tensor([0.05, -0.03, 1.2])
    c2w = torch.eye(4)[:3]          # OpenGL: camera looks down -z, which is towards the origin from +z
    c2w[:, 3] = cam
    fx = fy = 70.0
    cx, cy = W / 2, H / 2
    # z-depth of the sphere along every pixel ray (OpenCV camera after the diag(1,-1,-1) flip), 0 where the ray misses
    v, uu = torch.meshgrid(torch.arange(H) + 0.5, torch.arange(W) + 0.5, indexing="ij")
    d_cv = torch.stack([(uu - cx) / fx, (v - cy) / fy, torch.ones_like(uu)], dim=-1)     # z = 1
    R = c2w[:, :3] @ torch.diag(torch.tensor([1.0, -1.0, -1.0]))
    d_w = d_cv @ R.T
    b = (d_w * cam).sum(-1)
    a = (d_w * d_w).sum(-1)
    c = (cam * cam).sum() - 0.3 ** 2
    disc = b * b - a * c
    z = torch.where(disc > 0, (-b - disc.clamp(min=0).sqrt()) / a, torch.zeros_like(a))
    z = torch.where(z > 0, z + 0.003 * torch.randn(H, W, generator=g), torch.zeros_like(z))
    depth = z.clamp(min=0)[..., None].float()
    rgb = torch.rand(H, W, 3, generator=g)
    return dict(means=means, log_scales=log_scales, quats=quats, opacities=opac, normals=normals, c2w=c2w, fx=fx, fy=fy,
                cx=cx, cy=cy, H=H, W=W, depth=depth, rgb=rgb)


def main():
    sys.path.insert(0, str(ROOT))
    from tests import stubs

    stubs.REF = REF
    stubs.install()
    sys.modules["dn_splatter"].__path__ = [str(REF / "dn_splatter")]
    import dn_splatter.dn_model as ref_model
    from nerfstudio.cameras.cameras import Cameras

    s = scene()
    camera = Cameras(s["c2w"][None], s["fx"], s["fy"], s["cx"], s["cy"], s["W"], s["H"])
    fake_self = types.SimpleNamespace(
        means=s["means"], scales=s["log_scales"], quats=s["quats"], opacities=s["opacities"], normals=s["normals"],
        config=types.SimpleNamespace(knn_to_track=16), device=torch.device("cpu"),
        get_outputs=lambda camera: {"depth": s["depth"], "rgb": s["rgb"]})
    real_sample = random.sample
    random.sample = lambda population, k: list(population)[:k]
    try:
        out = ref_model.DNSplatterModel.compute_level_surface_points(fake_self, camera, num_samples=10 ** 9)
    finally:
        random.sample = real_sample
    arrays = {k: (v.numpy() if isinstance(v, torch.Tensor) else np.asarray(v)) for k, v in s.items()}
    for level, o in out.items():
        for key, t in o.items():
            arrays[f"L{level}_{key}"] = t.numpy()
    path = ROOT / "tests" / "golden" / "level_set.npz"
    np.savez_compressed(path, levels=np.asarray(list(out), dtype=np.float64), **arrays)
    print(path, path.stat().st_size, "bytes;", {lv: tuple(o["points"].shape) for lv, o in out.items()},
          "valid depth pixels:", int((s["depth"] > 0).sum()))


if __name__ == "__main__":
    main()
