"""f2: the level-set search of `compute_level_surface_points` (dn_model.py:1705-1946).
Golden: tests/golden/level_set.npz = the reference's own function run unmodified on the CPU (oracle/make_golden_level_set.py).
CPU: the oracle restatement reproduces it; GPU: the fused kernel and the host mirror reproduce both."""
import types
from pathlib import Path

import numpy as np
import pytest
import torch

from oracle import knn_ref, level_set_ref

GOLD = Path(__file__).resolve().parent / "golden" / "level_set.npz"


def _gold():
    g = {k: v for k, v in np.load(GOLD).items()}
    t = {k: torch.from_numpy(g[k]) for k in ("means", "log_scales", "quats", "opacities", "normals", "c2w", "depth", "rgb")}
    return g, t


def _points(g, t, device="cpu"):
    """The back-projected valid-depth points, their colours and the camera position (dn_model.py:1728-1758)."""
    from fusionsense_b200.level_set import backproject_depth

    H, W = int(g["H"]), int(g["W"])
    c2w = (t["c2w"] @ torch.diag(torch.tensor([1.0, -1.0, -1.0, 1.0]))).to(device)
    depth = t["depth"].to(device)
    pts = backproject_depth(depth, float(g["fx"]), float(g["fy"]), float(g["cx"]), float(g["cy"]), W, H, c2w)
    keep = ~(depth <= 0.0).reshape(-1)
    return pts[keep].contiguous(), t["rgb"].to(device).reshape(-1, 3)[keep], t["c2w"][:, 3].clone()


def test_oracle_level_set_matches_reference_golden():
    g, t = _gold()
    pts, cols, cam = _points(g, t)
    closest = torch.from_numpy(knn_ref.knn_sk_ref(t["means"].numpy(), pts.numpy(), 16))
    levels = [float(v) for v in g["levels"]]
    tt, valid, dens, _ = level_set_ref.level_crossings_ref(pts, cam, closest, t["means"], t["log_scales"], t["quats"],
                                                         t["opacities"], levels)
    dirs = torch.nn.functional.normalize(pts - cam[None], dim=-1)
    for li, lv in enumerate(levels):
        keep = valid[li]
        want = g[f"L{lv}_points"]
        assert int(keep.sum()) == len(want) > 100
        got = pts[keep] + tt[li][keep][:, None] * dirs[keep]
        np.testing.assert_allclose(got.numpy(), want, rtol=0, atol=2e-6)
        assert np.array_equal(cols[keep].numpy(), g[f"L{lv}_colors"])
        assert np.array_equal(t["normals"][closest[keep][:, 0]].numpy(), g[f"L{lv}_normals"])


DEV = "cuda"


@pytest.mark.gpu
def test_level_crossings_kernel_matches_oracle_and_reference_golden():
    from fusionsense_b200.knn import knn_sk
    from fusionsense_b200.level_set import level_crossings

    g, t = _gold()
    pts, cols, cam = _points(g, t)
    d = {k: v.to(DEV) for k, v in t.items()}
    levels = [float(v) for v in g["levels"]]
    closest = knn_sk(d["means"], pts.to(DEV), 16)
    assert np.array_equal(closest.cpu().numpy(), knn_ref.knn_sk_ref(t["means"].numpy(), pts.numpy(), 16))
    tt, valid, std, dens = level_crossings(pts.to(DEV), cam.tolist(), closest, d["means"], d["log_scales"], d["quats"],
                                           d["opacities"], levels, return_densities=True)
    rt, rvalid, rdens, rstd = level_set_ref.level_crossings_ref(pts, cam, closest.cpu(), t["means"], t["log_scales"],
                                                                t["quats"], t["opacities"], levels)
    np.testing.assert_allclose(std.cpu().numpy(), rstd.numpy(), rtol=2e-5)
    np.testing.assert_allclose(dens.cpu().numpy(), rdens.numpy(), rtol=1e-4, atol=1e-7)
    dirs = torch.nn.functional.normalize(pts - cam[None], dim=-1)
    for li, lv in enumerate(levels):
        # a ray's verdict may differ only where a sample's density sits within rounding of the level
        differ = valid[li].cpu() != rvalid[li]
        near = ((rdens - lv).abs() < 1e-5 * max(lv, 1.0)).any(dim=-1)
        assert not bool((differ & ~near).any())
        both = valid[li].cpu() & rvalid[li]
        np.testing.assert_allclose(tt[li].cpu()[both].numpy(), rt[li][both].numpy(), rtol=1e-3, atol=1e-6)
        if not bool(differ.any()):
            got = pts[both] + tt[li].cpu()[both][:, None] * dirs[both]
            np.testing.assert_allclose(got.numpy(), g[f"L{lv}_points"], rtol=0, atol=5e-6)


@pytest.mark.gpu
def test_host_mirror_reproduces_the_reference_function_on_its_golden_scene():
    """`level_set.compute_level_surface_points(model, camera, ...)` with a stand-in model whose get_outputs returns the
    golden depth / colour image: same rows, in pixel order (random.sample replaced by "the first k"), as the reference's
    own function produced on the CPU."""
    import random

    from fusionsense_b200 import level_set

    g, t = _gold()
    d = {k: v.to(DEV) for k, v in t.items()}
    model = types.SimpleNamespace(means=d["means"], scales=d["log_scales"], quats=d["quats"], opacities=d["opacities"],
                                  normals=d["normals"], config=types.SimpleNamespace(knn_to_track=16),
                                  get_outputs=lambda camera: {"depth": d["depth"], "rgb": d["rgb"]})
    one = lambda v, dt: torch.tensor([[v]], dtype=dt, device=DEV)  # noqa: E731
    camera = types.SimpleNamespace(camera_to_worlds=d["c2w"][None], fx=one(float(g["fx"]), torch.float32),
                                   fy=one(float(g["fy"]), torch.float32), cx=one(float(g["cx"]), torch.float32),
                                   cy=one(float(g["cy"]), torch.float32), width=one(int(g["W"]), torch.int64),
                                   height=one(int(g["H"]), torch.int64))
    real = random.sample
    random.sample = lambda population, k: list(population)[:k]
    try:
        out = level_set.compute_level_surface_points(model, camera, num_samples=10 ** 9)
        few = level_set.compute_level_surface_points(model, camera, num_samples=50)
    finally:
        random.sample = real
    for lv in (float(v) for v in g["levels"]):
        want = g[f"L{lv}_points"]
        assert out[lv]["points"].shape == want.shape
        np.testing.assert_allclose(out[lv]["points"].cpu().numpy(), want, rtol=0, atol=5e-6)
        assert np.array_equal(out[lv]["colors"].cpu().numpy(), g[f"L{lv}_colors"])
        assert np.array_equal(out[lv]["normals"].cpu().numpy(), g[f"L{lv}_normals"])
        assert few[lv]["points"].shape == (50, 3)
    with pytest.raises(NotImplementedError):
        level_set.compute_level_surface_points(model, camera, 10, return_normal="analytical")
