"""CPU: pin the oracle's vectorised stages against brute-force restatements of SURVEY.md Appendix A.

The reference has no tests or golden vectors for this path (parity unpinned, see oracle/gsplat_ref.py), so the
oracle is at least pinned against independent straight-line loops that follow the published kernel semantics
(A.4 keys / sort / offsets, A.5 forward compositing and the hand-written backward recurrences).
"""
import math

import numpy as np
import pytest
import torch

from fusionsense_b200.synthetic import make_scene
from oracle import gsplat_ref as ref


def _small(n=300, W=64, H=48, C=2, seed=41, mult=6.0):
    sc = make_scene(n, W, H, n_views=max(C, 2), cfg_id=seed, fx=60.0)
    scales = torch.exp(sc.scales) * mult
    out = ref.fully_fused_projection(sc.means, sc.quats, scales, sc.viewmats[:C], sc.Ks[:C], W, H)
    return sc, scales, out


def test_isect_against_loops():
    W, H, C = 64, 48, 2
    sc, scales, (radii, m2, dep, con, _) = _small(W=W, H=H, C=C)
    tw, th = math.ceil(W / 16), math.ceil(H / 16)
    tiles, ids, flat = ref.isect_tiles(m2, radii, dep, 16, tw, th, sort=False)
    tb = (tw * th).bit_length()
    exp_ids, exp_flat, exp_tiles = [], [], np.zeros((C, radii.shape[1]), np.int32)
    for c in range(C):
        for n in range(radii.shape[1]):
            r = int(radii[c, n])
            if r <= 0:
                continue
            x, y = np.float32(m2[c, n, 0]), np.float32(m2[c, n, 1])
            tr = np.float32(r) / np.float32(16)
            x0 = min(max(0, int(np.floor(x / np.float32(16) - tr))), tw)
            x1 = min(max(0, int(np.ceil(x / np.float32(16) + tr))), tw)
            y0 = min(max(0, int(np.floor(y / np.float32(16) - tr))), th)
            y1 = min(max(0, int(np.ceil(y / np.float32(16) + tr))), th)
            exp_tiles[c, n] = (y1 - y0) * (x1 - x0)
            dbits = int(np.float32(dep[c, n]).view(np.int32))
            for i in range(y0, y1):
                for j in range(x0, x1):
                    exp_ids.append((c << (32 + tb)) | ((i * tw + j) << 32) | dbits)
                    exp_flat.append(c * radii.shape[1] + n)
    assert np.array_equal(tiles.numpy(), exp_tiles)
    assert ids.tolist() == exp_ids and flat.tolist() == exp_flat
    assert len(exp_ids) > 100
    # sort + offsets
    _, ids_s, flat_s = ref.isect_tiles(m2, radii, dep, 16, tw, th, sort=True)
    order = sorted(range(len(exp_ids)), key=lambda k: (exp_ids[k], k))
    assert ids_s.tolist() == [exp_ids[k] for k in order]
    assert flat_s.tolist() == [exp_flat[k] for k in order]
    offs = ref.isect_offset_encode(ids_s, C, tw, th).reshape(-1).tolist()
    lin = [((k >> 32) >> tb) * tw * th + ((k >> 32) & ((1 << tb) - 1)) for k in ids_s.tolist()]
    for t in range(C * tw * th):
        assert offs[t] == sum(1 for v in lin if v < t)


def _brute_forward_backward(m2, con, col, op, W, H, offs, flat, bg, v_out, v_alpha):
    """Per-pixel loops, forward then the CUDA-style backward recurrences of Appendix A.5 (float64)."""
    C, N = op.shape
    D = col.shape[-1]
    th, tw = offs.shape[1], offs.shape[2]
    n_is = len(flat)
    ofl = offs.reshape(-1).tolist() + [n_is]
    out = np.zeros((C, H, W, D)); alpha = np.zeros((C, H, W, 1)); last = np.zeros((C, H, W), np.int64)
    g_m2 = np.zeros((C * N, 2)); g_abs = np.zeros((C * N, 2)); g_con = np.zeros((C * N, 3))
    g_col = np.zeros((C * N, D)); g_op = np.zeros(C * N)
    m2f, conf, colf, opf = m2.reshape(-1, 2), con.reshape(-1, 3), col.reshape(-1, D), op.reshape(-1)
    for c in range(C):
        for i in range(H):
            for j in range(W):
                lin = (c * th + i // 16) * tw + j // 16
                s, e = ofl[lin], ofl[lin + 1]
                px, py = j + 0.5, i + 0.5
                T = 1.0; acc = np.zeros(D); cur = 0
                for k in range(s, e):
                    g = flat[k]
                    dx, dy = m2f[g, 0] - px, m2f[g, 1] - py
                    sigma = 0.5 * (conf[g, 0] * dx * dx + conf[g, 2] * dy * dy) + conf[g, 1] * dx * dy
                    a = min(0.999, opf[g] * math.exp(-sigma))
                    if sigma < 0 or a < 1 / 255:
                        continue
                    nT = T * (1 - a)
                    if nT <= 1e-4:
                        break
                    acc += a * T * colf[g]; cur = k; T = nT
                out[c, i, j] = acc + (T * bg[c] if bg is not None else 0)
                alpha[c, i, j, 0] = 1 - T
                last[c, i, j] = cur
                # backward
                T_final = T; buf = np.zeros(D); vo = v_out[c, i, j]; va = v_alpha[c, i, j, 0]
                for k in range(cur, s - 1, -1):
                    g = flat[k]
                    dx, dy = m2f[g, 0] - px, m2f[g, 1] - py
                    sigma = 0.5 * (conf[g, 0] * dx * dx + conf[g, 2] * dy * dy) + conf[g, 1] * dx * dy
                    vis = math.exp(-sigma)
                    a = min(0.999, opf[g] * vis)
                    if sigma < 0 or a < 1 / 255:
                        continue
                    ra = 1 / (1 - a); T *= ra; fac = a * T
                    g_col[g] += fac * vo
                    v_a = ((colf[g] * T - buf * ra) * vo).sum() + T_final * ra * va
                    if bg is not None:
                        v_a += -T_final * ra * (bg[c] * vo).sum()
                    if opf[g] * vis <= 0.999:
                        v_s = -opf[g] * vis * v_a
                        g_con[g] += [0.5 * v_s * dx * dx, v_s * dx * dy, 0.5 * v_s * dy * dy]
                        gx = v_s * (conf[g, 0] * dx + conf[g, 1] * dy); gy = v_s * (conf[g, 1] * dx + conf[g, 2] * dy)
                        g_m2[g] += [gx, gy]; g_abs[g] += [abs(gx), abs(gy)]
                        g_op[g] += vis * v_a
                    buf += colf[g] * fac
    return out, alpha, last, g_m2, g_con, g_col, g_op


@pytest.mark.parametrize("bg", [False, True])
def test_compositing_forward_backward_against_loops(bg):
    W, H, C, D = 40, 24, 1, 3
    sc, scales, (radii, m2, dep, con, _) = _small(n=250, W=W, H=H, C=C, mult=10.0)
    tw, th = math.ceil(W / 16), math.ceil(H / 16)
    _, ids, flat = ref.isect_tiles(m2, radii, dep, 16, tw, th)
    offs = ref.isect_offset_encode(ids, C, tw, th)
    g = torch.Generator().manual_seed(3)
    col = torch.rand(C, radii.shape[1], D, generator=g, dtype=torch.float64)
    op = torch.sigmoid(sc.opacities[:, 0].double() + 2.0)[None]  # opaque enough to hit the T <= 1e-4 stop
    bgs = torch.rand(C, D, generator=g, dtype=torch.float64) if bg else None
    v_out = torch.randn(C, H, W, D, generator=g, dtype=torch.float64)
    v_alpha = torch.randn(C, H, W, 1, generator=g, dtype=torch.float64)
    ins = [t.double().clone().requires_grad_(True) for t in (m2, con, col, op)]
    out, alpha, last = ref.rasterize_to_pixels(*ins, W, H, 16, offs, flat, backgrounds=bgs, return_last_ids=True)
    ((out * v_out).sum() + (alpha * v_alpha).sum()).backward()
    b = _brute_forward_backward(ins[0].detach().numpy(), ins[1].detach().numpy(), ins[2].detach().numpy(),
                                ins[3].detach().numpy(), W, H, offs.numpy(), flat.tolist(),
                                None if bgs is None else bgs.numpy(), v_out.numpy(), v_alpha.numpy())
    assert (alpha.detach().numpy() > 0.9999 - 1e-4).any(), "scene should saturate somewhere (stop rule exercised)"
    np.testing.assert_allclose(out.detach().numpy(), b[0], rtol=1e-10, atol=1e-12)
    np.testing.assert_allclose(alpha.detach().numpy(), b[1], rtol=1e-10, atol=1e-12)
    assert np.array_equal(last.numpy(), b[2])
    np.testing.assert_allclose(ins[0].grad.reshape(-1, 2).numpy(), b[3], rtol=1e-7, atol=1e-10)
    np.testing.assert_allclose(ins[1].grad.reshape(-1, 3).numpy(), b[4], rtol=1e-7, atol=1e-10)
    np.testing.assert_allclose(ins[2].grad.reshape(-1, D).numpy(), b[5], rtol=1e-7, atol=1e-10)
    np.testing.assert_allclose(ins[3].grad.reshape(-1).numpy(), b[6], rtol=1e-7, atol=1e-10)


def test_rasterization_meta_and_modes():
    W, H, C, n = 64, 48, 2, 300
    sc = make_scene(n, W, H, n_views=2, cfg_id=43, fx=60.0)
    colors = torch.cat([sc.features_dc[:, None], sc.features_rest], 1)
    kw = dict(means=sc.means, quats=sc.quats, scales=torch.exp(sc.scales) * 6, opacities=torch.sigmoid(sc.opacities[:, 0]),
              viewmats=sc.viewmats, Ks=sc.Ks, width=W, height=H)
    r, a, meta = ref.rasterization(colors=colors, sh_degree=3, render_mode="RGB+ED", **kw)
    assert r.shape == (C, H, W, 4) and a.shape == (C, H, W, 1)
    assert meta["radii"].dtype == torch.int32 and meta["isect_ids"].dtype == torch.int64
    assert meta["isect_offsets"].shape == (C, 3, 4) and meta["tiles_per_gauss"].shape == (C, n)
    assert meta["isect_ids"].numel() == int(meta["tiles_per_gauss"].sum())
    # expected depth lies between the nearest and farthest visible depth wherever something was hit
    hit = a[..., 0] > 0.5
    vis_depth = meta["depths"][meta["radii"] > 0]
    assert (r[..., 3][hit] >= vis_depth.min() - 1e-4).all() and (r[..., 3][hit] <= vis_depth.max() + 1e-4).all()
    rgb, _, _ = ref.rasterization(colors=colors, sh_degree=3, render_mode="RGB", **kw)
    torch.testing.assert_close(rgb, r[..., :3])
    d_only, _, _ = ref.rasterization(colors=colors, sh_degree=3, render_mode="ED", **kw)
    torch.testing.assert_close(d_only[..., 0], r[..., 3])
    # lower SH degree only uses the first (deg+1)^2 bases
    r1, _, _ = ref.rasterization(colors=colors, sh_degree=1, render_mode="RGB", **kw)
    r1b, _, _ = ref.rasterization(colors=colors[:, :4], sh_degree=1, render_mode="RGB", **kw)
    torch.testing.assert_close(r1, r1b)


def test_legacy_bbox_is_superset_and_usually_equal():
    W, H = 64, 48
    sc, scales, (radii, m2, dep, con, _) = _small(W=W, H=H, C=1)
    tw, th = 4, 3
    t_new, _, _ = ref.isect_tiles(m2, radii, dep, 16, tw, th)
    t_old, _, _ = ref.isect_tiles(m2, radii, dep, 16, tw, th, legacy_bbox=True)
    assert (t_old >= t_new).all() and torch.equal(t_old, t_new)
    # an exact-integer edge makes the legacy box one tile wider (SURVEY.md A.6 caveat)
    m = torch.tensor([[[16.0, 20.0]]]); r = torch.tensor([[16]], dtype=torch.int32); z = torch.tensor([[1.0]])
    a, _, _ = ref.isect_tiles(m, r, z, 16, tw, th)
    b, _, _ = ref.isect_tiles(m, r, z, 16, tw, th, legacy_bbox=True)
    assert int(a) == 2 * 3 and int(b) == 3 * 3
