"""f3: seed point cloud — back-projection + voxel down-sample.  CPU: the oracle against
hand-computable cases; GPU: the CUDA path against the oracle (oracle/seed_points_ref.py: generate_pcd.py's own statements
on the CPU; open3d's VoxelDownSample restated — parity unpinned, open3d is not installed here)."""
import numpy as np
import pytest
import torch

from oracle import seed_points_ref as ref


def _view(H=96, W=128, seed=0):
    g = torch.Generator().manual_seed(seed)
    color = torch.rand(3, H, W, generator=g)
    depth = 0.2 + 1.2 * torch.rand(H, W, generator=g)
    depth[torch.rand(H, W, generator=g) < 0.1] = 0.0           # holes
    depth[torch.rand(H, W, generator=g) < 0.05] = 7.5          # beyond the 5 m cut
    depth[0, 0], depth[0, 1] = 0.5, 5.0                        # exactly on the open interval ends: dropped
    a = torch.tensor(0.3 + 0.1 * seed)
    R = torch.tensor([[torch.cos(a), 0, torch.sin(a)], [0, 1, 0], [-torch.sin(a), 0, torch.cos(a)]])
    w2c = torch.eye(4)
    w2c[:3, :3] = R
    w2c[:3, 3] = torch.tensor([0.1, -0.2, 0.4])
    return color, depth, w2c


def test_oracle_backprojection_hand_case():
    color = torch.zeros(3, 2, 2)
    color[0] = torch.tensor([[0.1, 0.2], [0.3, 0.4]])
    depth = torch.tensor([[0.25, 1.0], [0.0, 2.0]])
    fore, back = ref.get_pointcloud_ref(color, depth, torch.eye(4), 2.0, 2.0, 0.5, 0.5)
    assert fore.shape == (1, 6) and back.shape == (2, 6)
    np.testing.assert_allclose(fore[0].numpy(), [(0 - 0.5) / 2 * 0.25, (0 - 0.5) / 2 * 0.25, 0.25, 0.1, 0, 0], rtol=1e-6)
    np.testing.assert_allclose(back[1].numpy(), [(1 - 0.5) / 2 * 2.0, (1 - 0.5) / 2 * 2.0, 2.0, 0.4, 0, 0], rtol=1e-6)


def test_oracle_voxel_down_sample_hand_case():
    rows = np.array([[0.00, 0.0, 0.0, 1, 0, 0], [0.004, 0.0, 0.0, 0, 1, 0],   # same voxel (min_bound - 0.01 origin)
                     [0.05, 0.0, 0.0, 0, 0, 1], [0.0, 0.03, 0.0, 1, 1, 1]], dtype=np.float32)
    out = ref.voxel_down_sample_ref(rows, 0.02)
    assert out.shape == (3, 6)
    np.testing.assert_allclose(out[0], [0.002, 0, 0, 0.5, 0.5, 0], atol=1e-7)   # voxel (0,0,0)
    np.testing.assert_allclose(out[1], [0.0, 0.03, 0, 1, 1, 1], atol=1e-7)      # voxel (0,1,0) before (2,0,0): x major
    np.testing.assert_allclose(out[2], [0.05, 0, 0, 0, 0, 1], atol=1e-7)


DEV = "cuda"


@pytest.mark.gpu
@pytest.mark.parametrize("seed,transform", [(0, True), (1, True), (2, False)])
def test_get_pointcloud_matches_reference_statements(seed, transform):
    from fusionsense_b200.seed_points import get_pointcloud

    color, depth, w2c = _view(seed=seed)
    fx, fy, cx, cy = 130.0, 128.0, 63.2, 47.9
    fore_r, back_r = ref.get_pointcloud_ref(color, depth, w2c, fx, fy, cx, cy, transform_pts=transform)
    fore, back = get_pointcloud(color.to(DEV), depth.to(DEV), w2c.to(DEV), torch.tensor(fx, device=DEV), fy, cx, cy,
                                transform_pts=transform)
    assert fore.shape == fore_r.shape and back.shape == back_r.shape and back.dtype == torch.float32
    # same pixels in the same order (colours are copied, so they identify the pixel exactly)
    assert torch.equal(fore[:, 3:].cpu(), fore_r[:, 3:]) and torch.equal(back[:, 3:].cpu(), back_r[:, 3:])
    # fp32: the reference's matmul accumulates in an unspecified order and its 4x4 inverse comes from another LU
    np.testing.assert_allclose(back[:, :3].cpu().numpy(), back_r[:, :3].numpy(), rtol=0, atol=1e-5)
    np.testing.assert_allclose(fore[:, :3].cpu().numpy(), fore_r[:, :3].numpy(), rtol=0, atol=1e-5)


@pytest.mark.gpu
def test_get_pointcloud_empty_ranges():
    from fusionsense_b200.seed_points import get_pointcloud

    color, depth, w2c = _view()
    depth = torch.full_like(depth, 9.0)
    fore, back = get_pointcloud(color.to(DEV), depth.to(DEV), w2c.to(DEV), 100.0, 100.0, 64.0, 48.0)
    assert fore.shape == (0, 6) and back.shape == (0, 6)


@pytest.mark.gpu
@pytest.mark.parametrize("n,voxel", [(1, 0.02), (5000, 0.02), (5000, 0.5), (20000, 0.003)])
def test_voxel_down_sample_bit_exact_against_oracle(n, voxel):
    from fusionsense_b200.seed_points import voxel_down_sample

    g = torch.Generator().manual_seed(n)
    rows = torch.cat([torch.randn(n, 3, generator=g) * 0.3, torch.rand(n, 3, generator=g)], dim=1)
    if n > 100:
        rows[50:60] = rows[49]  # coincident points
    want = ref.voxel_down_sample_ref(rows.numpy(), voxel)
    got = voxel_down_sample(rows.to(DEV), voxel)
    assert got.dtype == torch.float64 and got.shape == want.shape
    assert np.array_equal(got.cpu().numpy(), want)  # same voxels, same order, same fp64 sums
    # xyz-only rows (a cloud without colours)
    got3 = voxel_down_sample(rows[:, :3].contiguous().to(DEV), voxel)
    assert np.array_equal(got3.cpu().numpy(), want[:, :3])


@pytest.mark.gpu
def test_voxel_down_sample_refusals_and_empty():
    from fusionsense_b200._abi import FsbError
    from fusionsense_b200.seed_points import voxel_down_sample

    rows = torch.rand(100, 6, device=DEV)
    with pytest.raises(ValueError):
        voxel_down_sample(rows, 0.0)
    with pytest.raises(FsbError, match="too small"):
        voxel_down_sample(rows * 1000.0, 1e-5)
    assert voxel_down_sample(rows[:0], 0.02).shape == (0, 6)
    with pytest.raises(RuntimeError):
        voxel_down_sample(rows.cpu(), 0.02)


@pytest.mark.gpu
def test_merged_background_cloud_matches_per_view_oracle():
    from fusionsense_b200.seed_points import merged_background_cloud

    views = [_view(seed=s) for s in range(3)]
    fx, fy, cx, cy = 130.0, 128.0, 63.2, 47.9
    got = merged_background_cloud([(c.to(DEV), d.to(DEV), w.to(DEV)) for c, d, w in views], fx, fy, cx, cy, 0.02)
    parts = []
    for c, d, w in views:
        _, back = ref.get_pointcloud_ref(c, d, w, fx, fy, cx, cy)
        parts.append(ref.voxel_down_sample_ref(back.numpy(), 0.02))
    want = np.concatenate(parts)
    # the back-projected coordinates differ in the last fp32 bits (matmul order), which can move a point across a voxel
    # face: compare the clouds as sets of voxels with a small allowance, and the colour / position sums globally
    assert abs(len(got) - len(want)) <= max(2, len(want) // 500)
    np.testing.assert_allclose(got.cpu().numpy().mean(0), want.mean(0), rtol=2e-3, atol=1e-4)
