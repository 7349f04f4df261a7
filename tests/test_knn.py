"""f2: exact KNN and the density field — oracle vs the reference's own outputs (CPU), CUDA path
vs both (GPU).  Golden: tests/golden/knn_sk.npz = dn_splatter/utils/knn.py::knn_sk (sklearn) and
DNSplatterModel.get_density run unmodified (oracle/make_golden_knn.py)."""
from pathlib import Path

import numpy as np
import pytest
import torch

from oracle import knn_ref

GOLD = Path(__file__).resolve().parent / "golden" / "knn_sk.npz"


def _gold():
    return {k: v for k, v in np.load(GOLD).items()}


def _same_neighbours(idx, ref_idx, x, y, tied=(5, 6)):
    """Index lists agree wherever the distances are distinct.  Where two candidates tie the order is an implementation
    detail of sklearn's heap, so there the DISTANCES must agree position by position and the row must hold one of the
    planted coincident points (`tied`: rows 5 and 6 of the golden cloud are the same point)."""
    idx, ref_idx = np.asarray(idx), np.asarray(ref_idx)
    assert idx.shape == ref_idx.shape
    bad = np.nonzero((idx != ref_idx).any(axis=1))[0]
    x64, y64 = x.astype(np.float64), y.astype(np.float64)
    for r in bad:
        d_a = np.sqrt(((x64[idx[r]] - y64[r]) ** 2).sum(-1))
        d_b = np.sqrt(((x64[ref_idx[r]] - y64[r]) ** 2).sum(-1))
        assert np.array_equal(d_a, d_b), (r, d_a, d_b)
        cols = idx[r] != ref_idx[r]
        assert set(idx[r][cols]) | set(ref_idx[r][cols]) <= set(tied), (r, idx[r], ref_idx[r])
    return len(bad)


def test_oracle_knn_matches_reference_sklearn_golden():
    g = _gold()
    k = int(g["k"])
    ties = _same_neighbours(knn_ref.knn_sk_ref(g["x"], g["x"], k), g["self_knn"], g["x"], g["x"])
    ties += _same_neighbours(knn_ref.knn_sk_ref(g["x"], g["y"], k), g["query_knn"], g["x"], g["y"])
    ties += _same_neighbours(knn_ref.knn_sk_ref(g["x"][:40], g["x"][:40], 3), g["small_knn"], g["x"][:40], g["x"][:40])
    assert 0 < ties < 40  # rows that see the planted duplicate pair, and nothing else


def test_oracle_density_matches_reference_golden():
    g = _gold()
    t = {k: torch.from_numpy(g[k]) for k in ("x", "y", "log_scales", "quats", "opacities")}
    d = knn_ref.get_density_ref(t["y"], torch.from_numpy(g["query_knn"]), t["x"], t["log_scales"], t["quats"],
                                t["opacities"])
    np.testing.assert_allclose(d.numpy(), g["density_query"], rtol=1e-6, atol=0)
    d = knn_ref.get_density_ref(t["x"][:2000], torch.from_numpy(g["self_knn"][:2000]), t["x"], t["log_scales"],
                                t["quats"], t["opacities"])
    np.testing.assert_allclose(d.numpy(), g["density_self"], rtol=1e-6, atol=0)


def test_kernel_algorithm_walkthrough_is_exact_on_the_golden_cloud():
    """tests/knn_emulation.py restates csrc/knn.cu's grid build and box-growth query in numpy; on the golden cloud
    (dense object, sparse shell, far outliers, a duplicate) its answers equal the brute-force oracle's, every query
    terminates on the grid, and none needs more than the kernel's default step budget."""
    from tests import knn_emulation as em

    g = _gold()
    x, y = g["x"], g["y"]
    ix = em.build(x)
    ref_idx, ref_dist = knn_ref.knn_full_ref(x, y[:120], 17)
    most = 0
    for r in range(120):
        best, steps, ok = em.query_one(ix, y[r], 17)
        assert ok
        assert [b[1] for b in best] == list(ref_idx[r])
        assert np.array_equal(np.sqrt([b[0] for b in best]), ref_dist[r])
        most = max(most, steps)
    assert most <= 96
    # a step budget of 0 visits the query's own cell only: (nearly) nobody is settled, nothing wrong is returned
    settled = [em.query_one(ix, y[r], 17, max_steps=0)[2] for r in range(40)]
    assert sum(settled) <= 4


# ---- CUDA path ----------------------------------------------------------------------------------------------------
DEV = "cuda"


@pytest.mark.gpu
def test_knn_sk_matches_reference_golden():
    from fusionsense_b200.knn import KnnIndex, knn_sk

    g = _gold()
    k = int(g["k"])
    x, y = torch.from_numpy(g["x"]).to(DEV), torch.from_numpy(g["y"]).to(DEV)
    out = knn_sk(x, x, k)
    assert out.dtype == torch.int64 and out.shape == (len(x), k) and out.is_cuda
    ties = _same_neighbours(out.cpu().numpy(), g["self_knn"], g["x"], g["x"])
    ties += _same_neighbours(knn_sk(x, y, k).cpu().numpy(), g["query_knn"], g["x"], g["y"])
    ties += _same_neighbours(knn_sk(x[:40].contiguous(), x[:40].contiguous(), 3).cpu().numpy(), g["small_knn"],
                             g["x"][:40], g["x"][:40])
    assert ties < 40
    # against the oracle the order is fully specified (distance, then index): bit-exact indices and distances
    index = KnnIndex(x)
    idx, dist = index.query(y, k + 1, return_distances=True)
    ref_idx, ref_dist = knn_ref.knn_full_ref(g["x"], g["y"], k + 1)
    assert np.array_equal(idx.cpu().numpy(), ref_idx)
    assert np.array_equal(dist.cpu().numpy(), ref_dist)
    # even the far outliers of the golden cloud settle on the grid (open-ended border cells): nobody needs the fallback
    assert int(index.last_unresolved) == 0


@pytest.mark.gpu
@pytest.mark.parametrize("n,k", [(1, 1), (2, 2), (33, 33), (1000, 5), (50000, 17)])
def test_knn_exact_against_oracle_on_random_clouds(n, k):
    from fusionsense_b200.knn import KnnIndex

    g = torch.Generator().manual_seed(n + k)
    x = torch.randn(n, 3, generator=g)
    x[: n // 3, 2] = 0.25            # a flat slab: degenerate extent on one axis for a third of the cloud
    if n >= 1000:
        x[10:20] = x[9]             # eleven coincident points
    ny = min(n, 700)
    y = torch.cat([x[torch.randint(0, n, (ny,), generator=g)], 3.0 * torch.randn(50, 3, generator=g)])
    index = KnnIndex(x.to(DEV))
    idx, dist = index.query(y.to(DEV), k, return_distances=True)
    ref_idx, ref_dist = knn_ref.knn_full_ref(x.numpy(), y.numpy(), k)
    assert np.array_equal(idx.cpu().numpy(), ref_idx)
    assert np.array_equal(dist.cpu().numpy(), ref_dist)
    # y is x: the cloud's own cell order is reused; the first neighbour of every point is a point at distance 0
    idx2, dist2 = index.query(index.x, k, return_distances=True)
    sub = torch.randint(0, n, (min(n, 300),), generator=g)
    ref_idx2, ref_dist2 = knn_ref.knn_full_ref(x.numpy(), x[sub].numpy(), k)
    assert np.array_equal(idx2[sub.to(DEV)].cpu().numpy(), ref_idx2)
    assert np.array_equal(dist2[sub.to(DEV)].cpu().numpy(), ref_dist2)
    assert float(dist2[:, 0].max()) == 0.0


@pytest.mark.gpu
def test_knn_edge_cases():
    from fusionsense_b200._abi import FsbError
    from fusionsense_b200.knn import KnnIndex, knn_sk

    x = torch.rand(100, 3, device=DEV)
    with pytest.raises(ValueError, match="n_neighbors <= n_samples_fit"):  # sklearn's refusal, same words
        knn_sk(x[:20].contiguous(), x[:20].contiguous(), 20)
    assert knn_sk(x, x[:0], 4).shape == (0, 4)
    with pytest.raises(ValueError):
        KnnIndex(x).query(x, 40)
    with pytest.raises((FsbError, AssertionError, RuntimeError)):
        knn_sk(x.cpu(), x.cpu(), 3)  # no CPU path
    # every point the same: all distances zero, indices 1..k (ties by index, first dropped)
    same = torch.ones(50, 3, device=DEV)
    assert torch.equal(knn_sk(same, same, 4), torch.arange(1, 5, device=DEV).expand(50, 4))
    # non-finite rows of x are never returned; a non-finite query gets -1
    x2 = x.clone()
    x2[7] = float("nan")
    x2[9, 1] = float("inf")
    q = x[:20].clone()
    q[3, 0] = float("nan")
    out = KnnIndex(x2).query(q, 5)
    assert not bool(((out == 7) | (out == 9)).any())
    assert bool((out[3] == -1).all()) and bool((out[[0, 1, 2, 4]] >= 0).all())
    ok = torch.ones(100, dtype=torch.bool)
    ok[[7, 9]] = False
    ref = knn_ref.knn_full_ref(x2[ok.to(DEV)].cpu().numpy(), q[:3].cpu().numpy(), 5)[0]
    remap = torch.nonzero(ok)[:, 0].numpy()
    assert np.array_equal(out[:3].cpu().numpy(), remap[ref])


@pytest.mark.gpu
def test_knn_every_query_through_the_brute_force_finish():
    """max_steps = 0 stops the grid walk after the query's own cell, which settles next to nothing: (almost) every query
    takes the fallback kernel and the answers stay exact."""
    from fusionsense_b200.knn import KnnIndex

    g = torch.Generator().manual_seed(5)
    x = torch.randn(3000, 3, generator=g)
    y = torch.randn(200, 3, generator=g)
    index = KnnIndex(x.to(DEV))
    idx, dist = index.query(y.to(DEV), 9, drop_first=1, return_distances=True, max_steps=0)
    assert int(index.last_unresolved) >= 150
    ref_idx, ref_dist = knn_ref.knn_full_ref(x.numpy(), y.numpy(), 9)
    assert np.array_equal(idx.cpu().numpy(), ref_idx[:, 1:])
    assert np.array_equal(dist.cpu().numpy(), ref_dist[:, 1:])


@pytest.mark.gpu
def test_gaussian_density_matches_reference_golden():
    from fusionsense_b200.knn import gaussian_density, knn_sk

    g = _gold()
    t = {k: torch.from_numpy(g[k]).to(DEV) for k in ("x", "y", "log_scales", "quats", "opacities")}
    closest = knn_sk(t["x"], t["y"], int(g["k"]))
    d = gaussian_density(t["y"], closest, t["x"], t["log_scales"], t["quats"], t["opacities"])
    np.testing.assert_allclose(d.cpu().numpy(), g["density_query"], rtol=2e-5, atol=0)
    closest = torch.from_numpy(g["self_knn"][:2000]).to(DEV)
    d = gaussian_density(t["x"][:2000].contiguous(), closest, t["x"], t["log_scales"], t["quats"], t["opacities"])
    np.testing.assert_allclose(d.cpu().numpy(), g["density_self"], rtol=2e-5, atol=0)


@pytest.mark.gpu
def test_knn_full_size_properties():
    """1M points (cfg4's Gaussian count), k = 16: properties that hold at any size — sorted distances, no self index
    after the drop, symmetric-difference-free agreement with the oracle on a sample of queries."""
    from fusionsense_b200.knn import KnnIndex

    g = torch.Generator().manual_seed(11)
    x = (torch.rand(1_000_000, 3, generator=g) * 2 - 1).to(DEV)
    index = KnnIndex(x)
    idx, dist = index.query(index.x, 17, drop_first=1, return_distances=True)
    assert idx.shape == (1_000_000, 16)
    assert bool((dist[:, 1:] >= dist[:, :-1]).all())
    assert not bool((idx == torch.arange(1_000_000, device=DEV)[:, None]).any())
    sub = torch.randint(0, 1_000_000, (64,), generator=g)
    ref_idx, ref_dist = knn_ref.knn_full_ref(x.cpu().numpy(), x[sub.to(DEV)].cpu().numpy(), 17, block=8)
    assert np.array_equal(idx[sub.to(DEV)].cpu().numpy(), ref_idx[:, 1:])
    assert np.array_equal(dist[sub.to(DEV)].cpu().numpy(), ref_dist[:, 1:])


@pytest.mark.parametrize("case", ["coincident", "slab", "line", "tiny"])
def test_kernel_algorithm_walkthrough_on_degenerate_clouds(case):
    """Quantile edges repeat when many points share a coordinate (empty zero-width cells); the box growth must still
    terminate with the exact answer (numpy walk-through of the kernel's algorithm, tests/knn_emulation.py)."""
    from tests import knn_emulation as em

    rng = np.random.default_rng(3)
    if case == "coincident":
        x = np.ones((60, 3), dtype=np.float32)
    elif case == "slab":
        x = rng.standard_normal((800, 3)).astype(np.float32)
        x[:500, 2] = 0.25
    elif case == "line":
        x = np.zeros((300, 3), dtype=np.float32)
        x[:, 0] = np.linspace(-1, 1, 300, dtype=np.float32)
    else:
        x = rng.standard_normal((5, 3)).astype(np.float32)
    k = min(5, len(x))
    ix = em.build(x)
    queries = np.concatenate([x[:25], rng.standard_normal((10, 3)).astype(np.float32) * 3])
    ref_idx, ref_dist = knn_ref.knn_full_ref(x, queries, k)
    for r, q in enumerate(queries):
        best, _, ok = em.query_one(ix, q, k, max_steps=10_000)
        assert ok
        assert [b[1] for b in best] == list(ref_idx[r]), (case, r)
        assert np.array_equal(np.sqrt([b[0] for b in best]), ref_dist[r])
