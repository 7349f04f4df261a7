"""a13 normal_from_depth_image: oracle pinned to the reference golden (CPU), CUDA kernel vs both (GPU)."""
import numpy as np
import pytest

from oracle import normal_utils_ref as ref

from .golden_io import GOLDEN

TOL = 1e-4  # fp32 tolerance of BASELINE.json north_star; normals have unit magnitude


def _golden():
    return np.load(GOLDEN / "pseudo_normals.npz")


def _close(a, b, frac=1e-3):
    err = np.abs(a - b)
    assert (err > TOL).mean() <= frac, (float(err.max()), float((err > TOL).mean()))


def test_oracle_matches_reference_golden():
    z = _golden()
    fx, fy, cx, cy = z["intr"]
    H, W = z["depth"].shape[:2]
    _close(ref.normal_from_depth_image(z["depth"], fx, fy, cx, cy, (W, H), np.eye(4)), z["n_eye"], frac=0.0)
    _close(ref.normal_from_depth_image(z["depth"], fx, fy, cx, cy, (W, H), z["c2w"]), z["n_pose"], frac=0.0)
    _close(ref.pcd_to_normal(z["xyz"]), z["n_pcd"], frac=0.0)
    assert not z["n_eye"][0].any() and not z["n_eye"][:, 0].any() and not z["n_eye"][-1].any()


@pytest.mark.gpu
def test_kernel_matches_reference_golden():
    import torch

    from fusionsense_b200.utils.normal_utils import normal_from_depth_image, pcd_to_normal

    z = _golden()
    fx, fy, cx, cy = (float(v) for v in z["intr"])
    H, W = z["depth"].shape[:2]
    dev = torch.device("cuda")
    d = torch.from_numpy(z["depth"]).to(dev)
    n = normal_from_depth_image(d, fx, fy, cx, cy, (W, H), torch.eye(4, device=dev), dev)
    _close(n.cpu().numpy(), z["n_eye"], frac=0.0)
    n = normal_from_depth_image(d, fx, fy, cx, cy, (W, H), torch.from_numpy(z["c2w"]).to(dev), dev)
    _close(n.cpu().numpy(), z["n_pose"], frac=0.0)
    _close(pcd_to_normal(torch.from_numpy(z["xyz"]).to(dev)).cpu().numpy(), z["n_pcd"], frac=0.0)
    gt = (1 + n.new_tensor(z["n_eye"]) @ torch.diag(torch.tensor([1.0, -1.0, -1.0], device=dev))) / 2
    _close(gt.cpu().numpy(), z["gt_normal"], frac=0.0)


@pytest.mark.gpu
@pytest.mark.parametrize("hw", [(480, 640), (1080, 1920), (3, 3), (2, 5), (1, 1)])
def test_kernel_matches_oracle_full_size(hw):
    import torch

    from fusionsense_b200.utils.normal_utils import normal_from_depth_image

    H, W = hw
    g = torch.Generator().manual_seed(H * 7919 + W)
    yy, xx = torch.meshgrid(torch.arange(H, dtype=torch.float32), torch.arange(W, dtype=torch.float32), indexing="ij")
    depth = 0.5 + 0.1 * torch.sin(xx / 37.0) * torch.cos(yy / 23.0) + 0.002 * torch.rand(H, W, generator=g)
    fx = fy = 600.0 * W / 640.0
    cx, cy = W / 2.0, H / 2.0
    want = ref.normal_from_depth_image(depth.numpy(), fx, fy, cx, cy, (W, H), np.eye(4))
    got = normal_from_depth_image(depth.cuda()[..., None], fx, fy, cx, cy, (W, H), torch.eye(4).cuda(),
                                  torch.device("cuda")).cpu().numpy()
    assert got.shape == (H, W, 3)
    _close(got, want)
    # size-independent properties: unit length in the interior, zero border
    if H > 2 and W > 2:
        ln = np.linalg.norm(got[1:-1, 1:-1], axis=-1)
        assert np.all(np.abs(ln - 1) < 1e-5)
    assert not got[0].any() and not got[-1].any() and not got[:, 0].any() and not got[:, -1].any()


def test_cpu_tensors_are_refused():
    import torch

    from fusionsense_b200.utils.normal_utils import normal_from_depth_image

    with pytest.raises(RuntimeError):
        normal_from_depth_image(torch.ones(4, 4, 1), 1.0, 1.0, 2.0, 2.0, (4, 4), torch.eye(4), torch.device("cpu"))
