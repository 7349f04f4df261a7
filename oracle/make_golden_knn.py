"""Generate tests/golden/knn_sk.npz by running the UNMODIFIED reference code in this container:

* `/root/reference/dn_splatter/utils/knn.py::knn_sk` (sklearn on the CPU) on a seeded surface-like cloud, with x is y
  (the `recompute_knn` / `populate_modules` call, dn_model.py:183-189, :306-310) and with separate query samples
  (`get_closest_gaussians`, dn_model.py:1562-1572);
* `/root/reference/dn_splatter/dn_model.py::DNSplatterModel.get_density` (the unbound function, called with a stand-in
  `self` that carries the four parameter tensors it reads) on those neighbours.  dn_model.py is imported against the
  stub nerfstudio / torchmetrics packages of tests/stubs (neither is installed here) and this repository's gsplat shim,
  whose `quat_to_rotmat` is plain torch.

Run from the repo root:  python -m oracle.make_golden_knn
"""
import importlib.util
import types
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parent.parent
REF = Path("/root/reference")


def cloud(n, seed):
    g = torch.Generator().manual_seed(seed)
    # points near the surface of three ellipsoids plus a sparse shell and a few far outliers: the shape of a trained
    # FusionSense scene (dense object, sparse background, strays)
    u = torch.nn.functional.normalize(torch.randn(n, 3, generator=g), dim=-1)
    radii = torch.tensor([[0.10, 0.06, 0.05], [0.05, 0.05, 0.08], [0.03, 0.07, 0.03]])
    centre = torch.tensor([[0.0, 0.0, 0.0], [0.08, 0.02, 0.03], [-0.05, -0.04, 0.06]])
    which = torch.randint(0, 3, (n,), generator=g)
    pts = centre[which] + u * radii[which] + 0.002 * torch.randn(n, 3, generator=g)
    shell = torch.rand(n, generator=g) < 0.15
    pts[shell] = 2.0 * torch.nn.functional.normalize(torch.randn(int(shell.sum()), 3, generator=g), dim=-1)
    pts[:5] = torch.tensor([[40.0, 0, 0], [0, -35.0, 1], [3, 3, 60.0], [-50.0, -50.0, -50.0], [25.0, 25.0, 0.0]])
    pts[5] = pts[6]  # an exact duplicate: a distance tie at zero
    return pts.float()


def main():
    spec = importlib.util.spec_from_file_location("ref_knn", REF / "dn_splatter" / "utils" / "knn.py")
    ref_knn = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(ref_knn)

    x = cloud(6000, 20241011)
    g = torch.Generator().manual_seed(7)
    pick = torch.randint(0, len(x), (1500,), generator=g)
    y = x[pick] + 0.01 * torch.randn(1500, 3, generator=g)  # samples near the surface (ray samples of the level-set search)
    y[:3] = torch.tensor([[100.0, 0, 0], [0.0, 0.0, 0.0], [-7.0, 9.0, 2.0]])
    k = 16
    self_knn = ref_knn.knn_sk(x, x, k)
    query_knn = ref_knn.knn_sk(x, y, k)
    small_knn = ref_knn.knn_sk(x[:40], x[:40], 3)

    # ---- get_density, the reference's own code --------------------------------------------------------------------
    import sys

    sys.path.insert(0, str(ROOT))
    from tests import stubs

    stubs.REF = REF  # import dn_splatter straight from the read-only reference tree
    stubs.install()
    sys.modules["dn_splatter"].__path__ = [str(REF / "dn_splatter")]
    import dn_splatter.dn_model as ref_model

    n = len(x)
    log_scales = torch.log(0.004 * torch.exp(0.5 * torch.randn(n, 3, generator=g)))
    log_scales[:50] = -9.0  # below the 1e-3 clamp of scale_rot_to_inv_cov3d
    quats = torch.randn(n, 4, generator=g)
    opac = 1.5 * torch.randn(n, 1, generator=g) + 1.0
    fake_self = types.SimpleNamespace(means=x, scales=log_scales, quats=quats, opacities=opac)
    with torch.no_grad():
        dens = ref_model.DNSplatterModel.get_density(fake_self, y, closest_gaussians=query_knn)
        dens_self = ref_model.DNSplatterModel.get_density(fake_self, x[:2000], closest_gaussians=self_knn[:2000])
    out = ROOT / "tests" / "golden" / "knn_sk.npz"
    np.savez_compressed(out, x=x.numpy(), y=y.numpy(), k=np.int64(k), self_knn=self_knn.numpy(),
                        query_knn=query_knn.numpy(), small_knn=small_knn.numpy(), log_scales=log_scales.numpy(),
                        quats=quats.numpy(), opacities=opac.numpy(), density_query=dens.numpy(),
                        density_self=dens_self.numpy())
    print(out, out.stat().st_size, "bytes;", "dens>=1:", int((dens_self > 0.99).sum()), "clamped:",
          int((dens <= 1e-4).sum()))


if __name__ == "__main__":
    main()
