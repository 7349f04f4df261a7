"""Exact k-nearest-neighbour search and the local density field on the GPU (csrc/knn.cu) — SURVEY.md §8 f2.

Host mirror of the reference's `dn_splatter/utils/knn.py` and of `DNSplatterModel.get_density`:

* `knn_sk(x, y, k)`  — same name, arguments and result as dn_splatter/utils/knn.py:29-43 (`NearestNeighbors(k + 1)
  .fit(x).kneighbors(y)` on the CPU, first column dropped): int64 `[len(y), k]` indices into `x`, nearest first.
  The reference's call sites (dn_model.py:183-189 `recompute_knn`, :306-310 `populate_modules`, :1562-1572
  `get_closest_gaussians`) pass CUDA tensors and get a CUDA tensor back; here nothing leaves the device except one
  4-byte read (the number of finite points, for sklearn's `n_neighbors <= n_samples_fit` refusal).
* `fast_knn(x, y, k)` — dn_splatter/utils/knn.py:9-26 (torch_cluster.knn wrapper, same result convention).
* `gaussian_density(samples, knn, means, scales, quats, opacities)` — dn_model.py:1596-1634 as one kernel.

A maintainer's binding is one line: `dn_splatter.dn_model.knn_sk = fusionsense_b200.knn.knn_sk`.
There is no CPU path: CPU tensors raise.
"""
from __future__ import annotations

import math
from typing import Optional, Tuple

import torch
from torch import Tensor

from ._abi import check, lib, ptr
from .ops import _f32c, _req_cuda, _stream, radix_sort_keys, radix_sort_pairs

MAX_K = 32            # k + 1 <= 33 candidates per query live in the query kernel's local list
TARGET_PER_CELL = 1.0
MAX_G = 320           # cells per axis
MAX_SAMPLE = 1 << 18  # points whose coordinates define the per-axis quantile edges
MAX_STEPS = 96        # growth steps of a query's cell box before it is handed to the brute-force finish


class KnnIndex:
    """Quantile-grid index over a fixed cloud `x` [N, 3] (float32, CUDA).  Build once, query many times."""

    def __init__(self, x: Tensor, target_per_cell: float = TARGET_PER_CELL):
        _req_cuda(x)
        assert x.dim() == 2 and x.shape[1] == 3, x.shape
        self.x = _f32c(x.detach())
        n = self.x.shape[0]
        assert 0 < n < 2 ** 30, n
        dev = self.x.device
        g = int(min(max(math.ceil((n / target_per_cell) ** (1.0 / 3.0)), 1), MAX_G))
        self.g = g
        self.n_cells = g ** 3
        # per-axis quantile edges from a strided sample
        S = min(n, MAX_SAMPLE)
        akeys = torch.empty((3 * S,), dtype=torch.int64, device=dev)
        check(lib.fsb_knn_axis_keys(S, n // S, ptr(self.x), ptr(akeys), _stream()), "fsb_knn_axis_keys")
        akeys, _ = radix_sort_keys(akeys, 0, 34, want_low32=False)
        self.edges = torch.empty((3, g + 1), dtype=torch.float32, device=dev)
        check(lib.fsb_knn_edges(S, ptr(akeys), g, ptr(self.edges), _stream()), "fsb_knn_edges")
        self._bits = max(1, int(self.n_cells).bit_length())
        bad = torch.zeros((1,), dtype=torch.int32, device=dev)
        keys, vals = self._cells(self.x, bad)
        keys, vals = radix_sort_pairs(keys, vals, self._bits)
        self.order = vals  # the cloud's own indices in cell order: the processing order when y is x
        self.cell_start = torch.empty((self.n_cells + 1,), dtype=torch.int32, device=dev)
        self.sorted_pts = torch.empty((n, 4), dtype=torch.float32, device=dev)
        check(lib.fsb_knn_build(n, ptr(keys), ptr(vals), ptr(self.x), g, ptr(self.cell_start), ptr(self.sorted_pts),
                                _stream()), "fsb_knn_build")
        self.n_finite = n - int(bad)  # the one host read (4 bytes): sklearn's n_samples_fit check needs it
        if self.n_finite < 1:
            raise ValueError("knn: the cloud holds no finite point")

    def _cells(self, pts: Tensor, n_nonfinite: Optional[Tensor] = None) -> Tuple[Tensor, Tensor]:
        n = pts.shape[0]
        keys = torch.empty((n,), dtype=torch.int64, device=pts.device)
        vals = torch.empty((n,), dtype=torch.int32, device=pts.device)
        check(lib.fsb_knn_cells(n, ptr(pts), self.g, ptr(self.edges), ptr(keys), ptr(vals), ptr(n_nonfinite),
                                _stream()), "fsb_knn_cells")
        return keys, vals

    def query(self, y: Tensor, k: int, drop_first: int = 0, return_distances: bool = False,
              max_steps: int = MAX_STEPS):
        """Indices [len(y), k - drop_first] (int64) of the k nearest points of `x` for every row of `y`, nearest first,
        ties by index, without the first `drop_first`; optionally the float64 distances as well."""
        _req_cuda(y)
        assert y.dim() == 2 and y.shape[1] == 3, y.shape
        if k > self.n_finite:
            # sklearn: "Expected n_neighbors <= n_samples_fit"
            raise ValueError(f"Expected n_neighbors <= n_samples_fit, but n_neighbors = {k}, n_samples_fit = "
                             f"{self.n_finite}")
        if not 1 <= k <= MAX_K + 1:
            raise ValueError(f"knn: k = {k} outside [1, {MAX_K + 1}] (the reference tracks knn_to_track = 16)")
        same = y.data_ptr() == self.x.data_ptr() and y.shape == self.x.shape and y.dtype == torch.float32
        yq = self.x if same else _f32c(y.detach())
        ny = yq.shape[0]
        dev = yq.device
        ko = k - drop_first
        out = torch.empty((ny, ko), dtype=torch.int64, device=dev)
        dist = torch.empty((ny, ko), dtype=torch.float64, device=dev) if return_distances else None
        if ny == 0:
            return (out, dist) if return_distances else out
        if same:
            order = self.order
        else:
            qk, qv = self._cells(yq)
            _, order = radix_sort_pairs(qk, qv, self._bits)
        unresolved = torch.empty((ny,), dtype=torch.int32, device=dev)
        n_unres = torch.empty((1,), dtype=torch.int32, device=dev)
        check(lib.fsb_knn_query(ny, ptr(yq), ptr(order), self.g, ptr(self.edges), ptr(self.cell_start),
                                ptr(self.sorted_pts), k, drop_first, max_steps, ptr(out), ptr(dist), ptr(unresolved),
                                ptr(n_unres), _stream()), "fsb_knn_query")
        check(lib.fsb_knn_brute(self.x.shape[0], ptr(self.x), ptr(yq), ptr(unresolved), ptr(n_unres), k, drop_first,
                                ptr(out), ptr(dist), _stream()), "fsb_knn_brute")
        self.last_unresolved = n_unres  # device counter (tests / tools read it)
        return (out, dist) if return_distances else out


def knn_sk(x: Tensor, y: Tensor, k: int) -> Tensor:
    """dn_splatter/utils/knn.py:29-43: the k + 1 nearest rows of `x` for every row of `y`, first one dropped."""
    return KnnIndex(x).query(y, k + 1, drop_first=1)


def fast_knn(x: Tensor, y: Tensor, k: int = 2) -> Tensor:
    """dn_splatter/utils/knn.py:9-26 (torch_cluster.knn(x, y, k + 1) reshaped, first column dropped)."""
    assert x.is_cuda and y.is_cuda and x.dim() == y.dim() == 2
    return knn_sk(x, y, k)


def gaussian_density(samples: Tensor, closest_gaussians: Tensor, means: Tensor, scales: Tensor, quats: Tensor,
                     opacities: Tensor, clamp_min: float = 1e-4) -> Tensor:
    """`DNSplatterModel.get_density` (dn_model.py:1596-1634) for given neighbours: samples [S,3], closest_gaussians
    [S,K] int64, the model's raw parameters (log-scales, wxyz quats, opacity logits [N,1]) -> clamped densities [S].
    `clamp_min=0` gives the unclamped copy inlined in compute_level_surface_points (dn_model.py:1806-1840).
    Forward only (the mesh-export path runs under no_grad)."""
    _req_cuda(samples, closest_gaussians, means, scales, quats, opacities)
    assert closest_gaussians.dtype == torch.int64 and closest_gaussians.dim() == 2
    S, K = closest_gaussians.shape
    assert samples.shape == (S, 3), samples.shape
    out = torch.empty((S,), dtype=torch.float32, device=samples.device)
    check(lib.fsb_gaussian_density(S, ptr(_f32c(samples.detach())), K, ptr(closest_gaussians.contiguous()),
                                   ptr(_f32c(means.detach())), ptr(_f32c(scales.detach())), ptr(_f32c(quats.detach())),
                                   ptr(_f32c(opacities.detach())), float(clamp_min), ptr(out), _stream()),
          "fsb_gaussian_density")
    return out
