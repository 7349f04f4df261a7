"""CPU: the per-Gaussian math the kernels inline (csrc/fs_math.cuh, built for the host) vs the oracle.

Forward values are compared with the fp32 oracle; the hand-derived backward formulas are compared with
fp64 autograd of the oracle's forward.
"""
import ctypes

import numpy as np
import pytest
import torch

from fusionsense_b200.synthetic import make_scene
from oracle import gsplat_ref as ref


def _p(a):
    return a.ctypes.data_as(ctypes.c_void_p)


def _cam16(viewmat, K):
    return np.concatenate([viewmat[:3].reshape(-1), [K[0, 0], K[1, 1], K[0, 2], K[1, 2]]]).astype(np.float32)


def _scene(n=4000, seed=1):
    sc = make_scene(n, 640, 480, n_views=3, cfg_id=seed)
    return sc, sc.means, sc.quats, torch.exp(sc.scales) * 8  # bigger footprints: exercise more branches


@pytest.mark.parametrize("cam", [0, 1, 2])
def test_projection_forward_matches_oracle(hostmath, cam):
    sc, means, quats, scales = _scene()
    n = means.shape[0]
    cam16 = _cam16(sc.viewmats[cam].numpy(), sc.Ks[cam].numpy())
    radii = np.zeros(n, np.int32); m2 = np.zeros((n, 2), np.float32); dep = np.zeros(n, np.float32)
    con = np.zeros((n, 3), np.float32); comp = np.zeros(n, np.float32)
    mn, qn, sn = means.numpy().copy(), quats.numpy().copy(), scales.numpy().copy()
    hostmath.hm_project_fwd(n, _p(cam16), _p(mn), _p(qn), _p(sn), 640, 480, ctypes.c_float(0.3),
                            ctypes.c_float(0.01), ctypes.c_float(1e10), ctypes.c_float(0.0), _p(radii), _p(m2),
                            _p(dep), _p(con), _p(comp))
    r, m, d, c, cp = ref.fully_fused_projection(means, quats, scales, sc.viewmats[cam:cam + 1], sc.Ks[cam:cam + 1],
                                                640, 480, calc_compensations=True)
    r, m, d, c, cp = r[0].numpy(), m[0].numpy(), d[0].numpy(), c[0].numpy(), cp[0].numpy()
    # depth feeds the sort key: bit exact
    both = (radii > 0) & (r > 0)
    assert (radii > 0).sum() > n // 10
    assert np.array_equal(dep[both].view(np.int32), d[both].view(np.int32))
    # visibility / radius may flip by fp32 rounding at a ceil() or cull boundary for a handful of Gaussians
    assert np.mean((radii > 0) != (r > 0)) < 2e-3
    assert np.mean(radii[both] != r[both]) < 5e-3
    np.testing.assert_allclose(m2[both], m[both], rtol=1e-4, atol=1e-3)
    np.testing.assert_allclose(con[both], c[both], rtol=2e-3, atol=1e-6)
    np.testing.assert_allclose(comp[both], cp[both], rtol=2e-3, atol=1e-5)


def test_projection_backward_matches_fp64_autograd(hostmath):
    sc, means, quats, scales = _scene(n=1500, seed=2)
    n = means.shape[0]
    cam = 1
    g = torch.Generator().manual_seed(5)
    v_m2 = torch.randn(n, 2, generator=g); v_d = torch.randn(n, generator=g)
    v_c = torch.randn(n, 3, generator=g); v_cp = torch.randn(n, generator=g)
    # fp64 autograd reference, and the same autograd in fp32 to calibrate what fp32 rounding alone does
    def autograd(dt):
        md, qd, sd = (t.detach().clone().to(dt).requires_grad_(True) for t in (means, quats, scales))
        vm = sc.viewmats[cam:cam + 1].to(dt).clone().requires_grad_(True)
        r, m, d, c, cp = ref.fully_fused_projection(md, qd, sd, vm, sc.Ks[cam:cam + 1].to(dt), 640, 480,
                                                    calc_compensations=True)
        loss = (m[0] * v_m2.to(dt)).sum() + (d[0] * v_d.to(dt)).sum() + (c[0] * v_c.to(dt)).sum() + (
            cp[0] * v_cp.to(dt)).sum()
        loss.backward()
        return r[0] > 0, md.grad.double().numpy(), qd.grad.double().numpy(), sd.grad.double().numpy(), vm.grad[0].double().numpy()

    vis, m64, q64, s64, gV = autograd(torch.float64)
    vis32, m32, q32, s32, _ = autograd(torch.float32)
    vis = vis & vis32
    cam16 = _cam16(sc.viewmats[cam].numpy(), sc.Ks[cam].numpy())
    gm = np.zeros((n, 3), np.float32); gq = np.zeros((n, 4), np.float32); gs = np.zeros((n, 3), np.float32)
    gR = np.zeros((n, 9), np.float32); gt = np.zeros((n, 3), np.float32)
    arrs = [means.numpy().copy(), quats.numpy().copy(), scales.numpy().copy(), v_m2.numpy().copy(),
            v_d.numpy().copy(), v_c.numpy().copy(), v_cp.numpy().copy()]
    hostmath.hm_project_bwd(n, _p(cam16), _p(arrs[0]), _p(arrs[1]), _p(arrs[2]), 640, 480, ctypes.c_float(0.3),
                            _p(arrs[3]), _p(arrs[4]), _p(arrs[5]), _p(arrs[6]), _p(gm), _p(gq), _p(gs), _p(gR),
                            _p(gt))
    v = vis.numpy()
    assert v.sum() > 100

    def errs(a, b):
        a, b = a[v], b[v]
        scale = np.abs(b).max(axis=-1, keepdims=True) + 1e-12
        e = np.abs(a - b) / scale
        return np.median(e), np.quantile(e, 0.99)

    # The hand-derived backward must be as close to fp64 autograd as fp32 autograd itself is (the chain
    # through 1/det^2 amplifies fp32 rounding; a wrong formula would be off by O(1), not by rounding).
    for name, mine, a32, a64 in (("means", gm, m32, m64), ("quats", gq, q32, q64), ("scales", gs, s32, s64)):
        med, q99 = errs(mine.astype(np.float64), a64)
        med32, q99_32 = errs(a32, a64)
        assert med <= 2 * med32 + 1e-6, (name, med, med32)
        assert q99 <= 2 * q99_32 + 1e-5, (name, q99, q99_32)
    # view-matrix gradient: sum over visible Gaussians
    R_sum = gR[v].astype(np.float64).sum(0).reshape(3, 3)
    t_sum = gt[v].astype(np.float64).sum(0)
    np.testing.assert_allclose(R_sum, gV[:3, :3], rtol=5e-3, atol=1e-3 * np.abs(gV[:3, :3]).max())
    np.testing.assert_allclose(t_sum, gV[:3, 3], rtol=5e-3, atol=1e-3 * np.abs(gV[:3, 3]).max())


@pytest.mark.parametrize("degree", [0, 1, 2, 3])
def test_sh_basis_and_gradient(hostmath, degree):
    g = torch.Generator().manual_seed(7)
    n = 512
    d = torch.randn(n, 3, generator=g)
    u = (d / d.norm(dim=-1, keepdim=True)).float()
    basis = np.zeros((n, 16), np.float32); dx = np.zeros((n, 16), np.float32)
    dy = np.zeros((n, 16), np.float32); dz = np.zeros((n, 16), np.float32)
    un = u.numpy().copy()
    hostmath.hm_sh_basis(n, degree, _p(un), _p(basis), _p(dx), _p(dy), _p(dz))
    nb = (degree + 1) ** 2
    # closed-form polynomials evaluated WITHOUT renormalising, so the Jacobian is the plain polynomial one
    ud = u.double().requires_grad_(True)
    x, y, z = ud.unbind(-1)
    B = ref.sh_bases(degree, ud)  # normalises inside; on the unit sphere the values agree
    np.testing.assert_allclose(basis[:, :nb], B.detach().numpy(), rtol=1e-5, atol=1e-6)
    # gradient check through the normalisation: project both onto the tangent plane
    for k in range(1, nb):  # basis 0 is a constant
        (gk,) = torch.autograd.grad(B[:, k].sum(), ud, retain_graph=True)
        mine = np.stack([dx[:, k], dy[:, k], dz[:, k]], -1).astype(np.float64)
        un64 = u.double().numpy()
        mine_t = mine - un64 * (mine * un64).sum(-1, keepdims=True)
        np.testing.assert_allclose(mine_t, gk.numpy(), rtol=1e-4, atol=2e-5)
