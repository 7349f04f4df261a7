"""GPU parity of voxel carving: libfsb200's fsb_vh_* kernels vs the reference golden (201^3, bit exact) and vs the
numpy oracle on other grids, plus slab sharding and the reference-facing VisualHull() entry point."""
import numpy as np
import pytest
import torch

from oracle import visual_hull_ref as vh_ref
from tests.golden_io import load_visual_hull_golden, write_visual_hull_capture

pytestmark = pytest.mark.gpu


def test_visual_hull_matches_reference_golden_bit_exact(tmp_path):
    from fusionsense_b200.visual_hull import VisualHull

    g = load_visual_hull_golden()
    path = write_visual_hull_capture(tmp_path / "cap", g)
    out = tmp_path / "out"
    pts = VisualHull(path, str(out), error=5)
    assert pts.dtype == np.float64 and pts.shape == g["points"].shape
    assert np.array_equal(pts, g["points"])  # same voxels, same order, same coordinates
    # the PLY on disk holds the same doubles
    raw = (out / "foreground_pcd.ply").read_bytes()
    body = raw[raw.index(b"end_header\n") + len(b"end_header\n"):]
    assert np.array_equal(np.frombuffer(body, dtype="<f8").reshape(-1, 3), g["points"])


@pytest.mark.parametrize("n_axis,world", [(64, 1), (97, 3), (128, 2)])
def test_votes_and_slabs_match_numpy_oracle(tmp_path, n_axis, world):
    from fusionsense_b200 import visual_hull as vh

    g = load_visual_hull_golden()
    path = write_visual_hull_capture(tmp_path, g)
    mats, centre, names = vh.read_hull_cameras(path)
    cams = vh_ref.cameras_from_transforms(path)
    assert np.array_equal(mats, cams.mats) and np.array_equal(centre, cams.camera_center) and names == cams.names
    masks = vh.read_masks(path, names)
    # grey-level masks: exercise the value/255 table, not only 0/255
    rng = np.random.default_rng(3)
    masks = np.where(masks > 0, rng.integers(1, 256, masks.shape, dtype=np.uint8), 0).astype(np.uint8)
    xs, ys, zs = vh.hull_grid(centre, half_extent=0.2, n_per_axis=n_axis)
    rx, ry, rz = vh_ref.grid_axes(cams.camera_center, half_extent=0.2, n_per_axis=n_axis)
    assert np.array_equal(xs, rx) and np.array_equal(ys, ry) and np.array_equal(zs, rz)
    votes_ref = vh_ref.project_votes(cams.mats, masks, xs, ys, zs)
    maxv_ref, iso_ref = vh_ref.threshold(votes_ref, 5)
    pts_ref = vh_ref.occupied_points(votes_ref, iso_ref, xs, ys, zs)

    carvers = [vh.HullCarver(mats, masks, xs, ys, zs, rank=r, world_size=world) for r in range(world)]
    maxv = max(c.vote() for c in carvers)
    votes = torch.cat([c.votes for c in carvers]).cpu().numpy()
    assert np.array_equal(votes, votes_ref)  # float64 sums in view order: bit exact
    assert maxv == maxv_ref and vh.iso_value(maxv, 5) == iso_ref
    parts = [c.extract(iso_ref, want_indices=True) for c in carvers]
    pts = torch.cat([p for p, _ in parts]).cpu().numpy()
    idx = torch.cat([i for _, i in parts]).cpu().numpy()
    assert np.array_equal(pts, pts_ref)
    assert np.array_equal(idx, np.nonzero(votes_ref > iso_ref)[0])


def test_out_of_frustum_voxels_follow_reference_clamp_rules(tmp_path):
    from fusionsense_b200 import visual_hull as vh

    g = load_visual_hull_golden()
    path = write_visual_hull_capture(tmp_path, g)
    mats, centre, names = vh.read_hull_cameras(path)
    masks = vh.read_masks(path, names)
    masks[:, 0, :] = 255  # first row / column lit: clamped projections pick these up
    masks[:, :, 0] = 255
    xs = np.linspace(-3, 3, 23); ys = np.linspace(-3, 3, 19); zs = np.linspace(3, -3, 17)
    # include the camera centres themselves (p_z == 0 -> division by zero / NaN path)
    c2w = g["c2w"]
    xs = np.concatenate([xs, c2w[:, 0, 3]]); ys = np.concatenate([ys, c2w[:, 1, 3]]); zs = np.concatenate([zs, c2w[:, 2, 3]])
    votes_ref = vh_ref.project_votes(mats, masks, xs, ys, zs)
    c = vh.HullCarver(mats, masks, xs, ys, zs)
    c.vote()
    assert np.array_equal(c.votes.cpu().numpy(), votes_ref)
