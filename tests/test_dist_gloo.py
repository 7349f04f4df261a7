"""CPU, world_size 2, gloo: the host-side sharding logic of the N>1 paths (no GPU needed).

  * visual hull: slab bounds, MAX all-reduce of the vote maximum, rank-ordered all-gather of the occupied points
    -> identical to the single-process result (the per-slab votes come from the numpy oracle here; on a GPU box the
    same functions are fed by the fsb_vh_* kernels, tests/test_gpu_visual_hull.py).
  * training: the flat gradient all-reduce bench.py uses keeps replicas identical.
"""
import os
import sys
from pathlib import Path

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = Path(__file__).resolve().parent.parent


def _hull_worker(rank, world, port, tmp, ret):
    sys.path.insert(0, str(ROOT))
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from fusionsense_b200 import visual_hull as vh
        from oracle import visual_hull_ref as ref

        mats, centre, names = vh.read_hull_cameras(tmp)
        masks = vh.read_masks(tmp, names)
        xs, ys, zs = vh.hull_grid(centre, half_extent=0.15, n_per_axis=41)
        z0, z1 = vh.slab_bounds(len(zs), rank, world)
        votes = ref.project_votes(mats, masks, xs, ys, zs[z0:z1])  # stands in for HullCarver.vote() on CPU
        maxv = vh.reduce_max(float(votes.max()) if votes.size else 0.0, "cpu")
        iso = vh.iso_value(maxv, 5)
        pts = torch.from_numpy(ref.occupied_points(votes, iso, xs, ys, zs[z0:z1]))
        allpts = vh.gather_slabs(pts)
        if rank == 0:
            np.save(ret, allpts.numpy())
            np.save(ret + ".meta.npy", np.array([maxv, iso]))
    finally:
        dist.destroy_process_group()


def test_hull_slab_sharding_world2(tmp_path):
    from fusionsense_b200 import visual_hull as vh
    from oracle import visual_hull_ref as ref
    from tests.golden_io import load_visual_hull_golden, write_visual_hull_capture

    g = load_visual_hull_golden()
    path = write_visual_hull_capture(tmp_path / "cap", g)
    ret = str(tmp_path / "ret.npy")
    mp.spawn(_hull_worker, args=(2, 29531, path, ret), nprocs=2, join=True)
    got = np.load(ret)
    maxv, iso = np.load(ret + ".meta.npy")
    cams = ref.cameras_from_transforms(path)
    masks = ref.load_masks(path, cams.names)
    xs, ys, zs = ref.grid_axes(cams.camera_center, half_extent=0.15, n_per_axis=41)
    votes = ref.project_votes(cams.mats, masks, xs, ys, zs)
    m_ref, iso_ref = ref.threshold(votes, 5)
    assert (maxv, iso) == (m_ref, iso_ref)
    assert np.array_equal(got, ref.occupied_points(votes, iso_ref, xs, ys, zs))
    assert got.shape[0] > 10


def test_slab_bounds_partition():
    from fusionsense_b200.visual_hull import slab_bounds

    for nz in (1, 7, 201, 512):
        for world in (1, 2, 3, 4, 8):
            b = [slab_bounds(nz, r, world) for r in range(world)]
            assert b[0][0] == 0 and b[-1][1] == nz
            assert all(b[i][1] == b[i + 1][0] for i in range(world - 1))
            sizes = [z1 - z0 for z0, z1 in b]
            assert max(sizes) - min(sizes) <= 1


def _grad_worker(rank, world, port, ret):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        torch.manual_seed(0)
        params = [torch.nn.Parameter(torch.randn(50, 3)), torch.nn.Parameter(torch.randn(50, 15, 3))]
        opt = torch.optim.Adam(params, lr=1e-2, eps=1e-15)
        for it in range(3):
            g = torch.Generator().manual_seed(100 * it + rank)  # every rank sees a different camera view
            for p in params:
                p.grad = torch.randn(p.shape, generator=g)
            flat = torch.cat([p.grad.reshape(-1) for p in params])
            dist.all_reduce(flat)
            flat.div_(world)
            o = 0
            for p in params:
                p.grad = flat[o:o + p.numel()].view_as(p)
                o += p.numel()
            opt.step()
        torch.save([p.detach().clone() for p in params], f"{ret}.{rank}")
    finally:
        dist.destroy_process_group()


def test_flat_gradient_allreduce_keeps_replicas_identical(tmp_path):
    ret = str(tmp_path / "params")
    mp.spawn(_grad_worker, args=(2, 29533, ret), nprocs=2, join=True)
    a, b = torch.load(f"{ret}.0"), torch.load(f"{ret}.1")
    for x, y in zip(a, b):
        assert torch.equal(x, y)


class _StatsModel:
    """the attributes sync_densify_stats reads from DNSplatterModel / DNSplatterStep"""

    def __init__(self, n, rank):
        g = torch.Generator().manual_seed(7 + rank)
        self.gauss_params = {"means": torch.zeros(n, 3)}
        self.num_points = n
        self.xys_grad_norm = torch.rand(n, generator=g)
        self.vis_counts = torch.ones(n) + torch.randint(0, 5, (n,), generator=g).float()
        self.max_2Dsize = torch.rand(n, generator=g) if rank == 0 else None  # rank 1 accumulated nothing yet


def _sync_worker(rank, world, port, ret):
    sys.path.insert(0, str(ROOT))
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from fusionsense_b200.dist import GradSync, split_generator, sync_densify_stats

        m = _StatsModel(64, rank)
        sync_densify_stats(m)
        torch.manual_seed(0)
        params = [torch.nn.Parameter(torch.randn(64, 3)), torch.nn.Parameter(torch.randn(64, 1))]
        g = torch.Generator().manual_seed(100 + rank)
        for p in params:
            p.grad = torch.randn(p.shape, generator=g)
        overflow = torch.tensor([1 if rank == 1 else 0], dtype=torch.int32)
        sync = GradSync()
        sync(params, overflow)
        draws = torch.randn(5, 3, generator=split_generator(3, 1200, "cpu"))
        torch.save({"grad": [p.grad.clone() for p in params], "overflow": overflow, "draws": draws,
                    "stats": (m.xys_grad_norm, m.vis_counts, m.max_2Dsize)}, f"{ret}.{rank}")
    finally:
        dist.destroy_process_group()


def test_grad_sync_and_densify_stats_world2(tmp_path):
    ret = str(tmp_path / "sync")
    mp.spawn(_sync_worker, args=(2, 29535, ret), nprocs=2, join=True)
    a, b = torch.load(f"{ret}.0"), torch.load(f"{ret}.1")
    for x, y in zip(a["grad"], b["grad"]):
        assert torch.equal(x, y)
    g0, g1 = torch.Generator().manual_seed(100), torch.Generator().manual_seed(101)
    want = torch.randn(64, 3, generator=g0) + torch.randn(64, 3, generator=g1)
    assert torch.allclose(a["grad"][0], want)
    assert int(a["overflow"]) == 1 and int(b["overflow"]) == 1  # any rank's overflow skips the step everywhere
    assert torch.equal(a["draws"], b["draws"])
    m0, m1 = _StatsModel(64, 0), _StatsModel(64, 1)
    for x, y in zip(a["stats"], b["stats"]):
        assert torch.equal(x, y)
    assert torch.allclose(a["stats"][0], m0.xys_grad_norm + m1.xys_grad_norm)
    assert torch.equal(a["stats"][1], m0.vis_counts + m1.vis_counts - 1.0)
    assert torch.equal(a["stats"][2], m0.max_2Dsize)


def test_shard_views_covers_global_batch():
    from fusionsense_b200.dist import shard_views

    for world in (1, 2, 4, 8):
        for step in range(5):
            got = sorted(v for r in range(world) for v in shard_views(step, r, world, n_views=1000))
            assert got == list(range(step * world, (step + 1) * world))
    assert shard_views(3, 1, 2, n_views=9) == [(3 * 2 + 1) % 9]
