"""GPU: static-capacity mode of the intersection pipeline (no host read of n_isects) and the CUDA-graph-captured
training iteration built on it, against the eager path."""
import pytest
import torch

from fusionsense_b200.synthetic import make_scene
from tests.parity import assert_close

pytestmark = pytest.mark.gpu
DEV = "cuda"


def _scene(n=20000, W=320, H=240, views=3):
    return make_scene(n, W, H, n_views=views, cfg_id=61, kind="bunny", fx=300.0)


def _render_inputs(sc, cam=0):
    return dict(means=sc.means, quats=sc.quats / sc.quats.norm(dim=-1, keepdim=True), scales=torch.exp(sc.scales),
                opacities=torch.sigmoid(sc.opacities).squeeze(-1),
                colors=torch.cat((sc.features_dc[:, None, :], sc.features_rest), dim=1),
                viewmats=sc.viewmats[cam:cam + 1], Ks=sc.Ks[cam:cam + 1], width=sc.width, height=sc.height,
                packed=False, render_mode="RGB+ED", sh_degree=3, absgrad=True)


@pytest.mark.parametrize("slack", [1, 4097, 300000])
def test_static_capacity_lists_are_bit_exact(slack):
    """Same keys, same sort order, same offsets, same image as the exact-size path; the tail of the buffers beyond
    the true count is never read."""
    from fusionsense_b200 import ops
    from fusionsense_b200.gsplat import rasterization

    sc = _scene().to(DEV)
    img0, a0, meta0 = rasterization(**_render_inputs(sc))
    n = meta0["flatten_ids"].numel()
    flag = torch.zeros(1, dtype=torch.int32, device=DEV)
    with ops.static_capacity(n + slack, flag) as st:
        img1, a1, meta1 = rasterization(**_render_inputs(sc))
        assert len(st.counts) == 1 and int(st.counts[0]) == n
    assert int(flag) == 0
    assert meta1["flatten_ids"].numel() == n + slack
    assert torch.equal(meta1["isect_ids"][:n], meta0["isect_ids"])
    assert torch.equal(meta1["flatten_ids"][:n], meta0["flatten_ids"])
    assert torch.equal(meta1["isect_offsets"], meta0["isect_offsets"])
    assert torch.equal(img1, img0) and torch.equal(a1, a0)


def test_static_capacity_backward_matches_and_overflow_is_flagged():
    from fusionsense_b200 import ops
    from fusionsense_b200.gsplat import rasterization

    sc = _scene().to(DEV)
    grads = []
    n = None
    for static in (False, True):
        inp = _render_inputs(sc)
        leaves = {k: inp[k].detach().clone().requires_grad_(True) for k in ("means", "quats", "scales", "opacities", "colors")}
        inp.update(leaves)
        flag = torch.zeros(1, dtype=torch.int32, device=DEV)
        if static:
            with ops.static_capacity(n + 12345, flag):
                img, alpha, meta = rasterization(**inp)
                ((img * img).sum() + alpha.sum()).backward()
            assert int(flag) == 0
        else:
            img, alpha, meta = rasterization(**inp)
            n = meta["flatten_ids"].numel()
            ((img * img).sum() + alpha.sum()).backward()
        grads.append({k: v.grad.clone() for k, v in leaves.items()})
    for k in grads[0]:
        assert_close(grads[1][k], grads[0][k], f"static.grad.{k}", tol=1e-5, outlier_frac=1e-4)
    # capacity too small: flagged, nothing out of bounds (compute-sanitizer clean by construction: writes clamp)
    flag = torch.zeros(1, dtype=torch.int32, device=DEV)
    with ops.static_capacity(max(1, n // 3), flag):
        img, alpha, meta = rasterization(**_render_inputs(sc))
    torch.cuda.synchronize()
    assert int(flag) == 1 and torch.isfinite(img).all()


@pytest.mark.parametrize("force_mismatch", [False, True])
def test_device_side_list_sharing_for_the_legacy_pass(force_mismatch):
    """Static-capacity mode: the legacy (0.1.x bbox) binning that follows a 1.0 binning of the same Gaussians either
    reuses its sorted lists (equal totals) or builds its own (a Gaussian whose bbox edge sits exactly on a tile
    boundary gains a tile under the legacy rule) -- decided on the device, bit exact against the legacy binning
    done from scratch in both cases."""
    from fusionsense_b200 import ops

    g = torch.Generator().manual_seed(5)
    N, W, H, ts = 5000, 320, 240, 16
    tw, th = W // ts, H // ts
    m2 = (torch.rand(1, N, 2, generator=g) * torch.tensor([W, H])).to(DEV)
    radii = torch.randint(0, 40, (1, N), generator=g, dtype=torch.int32).to(DEV)
    depths = (torch.rand(1, N, generator=g) * 5 + 0.1).to(DEV)
    if force_mismatch:
        m2[0, :50, 0] = 64.0  # (x + r) / 16 is an integer: ceil() keeps it, the legacy (int)(.. + 1) adds a tile
        radii[0, :50] = 32
    flag = torch.zeros(1, dtype=torch.int32, device=DEV)
    _, _, flat_ref, offs_ref = ops.isect_tiles(m2, radii, depths, ts, tw, th, legacy_bbox=True)
    _, _, flat_10, _ = ops.isect_tiles(m2, radii, depths, ts, tw, th, legacy_bbox=False)
    n_ref = flat_ref.numel()
    assert (flat_10.numel() != n_ref) == force_mismatch
    with ops.static_capacity(n_ref + 1000, flag) as st:
        _, _, first_flat, first_offs = ops.isect_tiles(m2, radii, depths, ts, tw, th, legacy_bbox=False)
        done = torch.cuda.Event()
        done.record()
        flat, offs = ops.isect_tiles_legacy_shared(m2, radii, depths, ts, tw, th, first_flat, first_offs,
                                                   lists_done=done)
        assert int(flat.n_dev) == n_ref and int(st.counts[-1]) == n_ref
    assert int(flag) == 0
    assert torch.equal(flat[:n_ref], flat_ref)
    assert torch.equal(offs, offs_ref)


def _pair(**kw):
    from fusionsense_b200.dn_step import DNSplatterStep, DNSplatterStepConfig

    sc = _scene(n=15000, W=256, H=192, views=3)
    eager = DNSplatterStep(sc, DNSplatterStepConfig(**kw), device=DEV, step=3000)
    graphed = DNSplatterStep(sc, DNSplatterStepConfig(**kw), device=DEV, step=3000)
    targets = {v: eager.render_targets(v) for v in range(3)}
    return eager, graphed, targets


def test_graphed_iterations_track_eager_iterations():
    from fusionsense_b200.graph_step import GraphedDNSplatterStep

    eager, graphed, targets = _pair()
    runner = GraphedDNSplatterStep(graphed, targets)
    order = [0, 1, 2, 1, 0, 2, 2, 0]
    losses_e, losses_g = [], []
    for v in order:
        losses_e.append(float(eager.train_iteration(v, targets[v])))
        runner.train_iteration(v)
        losses_g.append(runner.poll()["loss"])
    info = runner.poll()
    assert runner.captures == 1 and runner.replays == len(order) and info["overflowed_steps"] == 0
    # fused_passes (default): one union list serves both colour sets, there is no second binning to count
    assert info["n_isects"] > 0 and info["n_isects_normals"] in (0, info["n_isects"], info["n_isects"] + 1)
    assert graphed.step == eager.step
    for a, b in zip(losses_g, losses_e):
        assert a == pytest.approx(b, rel=2e-4)
    # float atomics make the two runs differ in the last bits of every gradient; after 8 Adam steps with
    # eps = 1e-15 a sign flip of a ~0 gradient moves a parameter by up to lr per step, hence the outlier budget
    for k in eager.gauss_params:
        assert_close(graphed.gauss_params[k].data, eager.gauss_params[k].data, f"graph.param.{k}", tol=2e-3,
                     outlier_frac=2e-2)
    for k in ("vis_counts", "max_2Dsize"):
        assert_close(getattr(graphed, k), getattr(eager, k), f"graph.stats.{k}", tol=1e-6, outlier_frac=1e-3)
    for name, opt in graphed.optimizers.items():
        p = opt.param_groups[0]["params"][0]
        assert float(opt.state[p]["step"]) == float(eager.optimizers[name].state[eager.gauss_params[name]]["step"])


def test_graphed_single_step_matches_eager_step_closely():
    """One step from identical state: same loss to fp32 rounding, same parameters after Adam."""
    from fusionsense_b200.graph_step import GraphedDNSplatterStep

    eager, graphed, targets = _pair()
    runner = GraphedDNSplatterStep(graphed, targets)
    le = float(eager.train_iteration(1, targets[1]))
    runner.train_iteration(1)
    assert runner.poll()["loss"] == pytest.approx(le, rel=1e-5)
    for k in eager.gauss_params:
        assert_close(graphed.gauss_params[k].data, eager.gauss_params[k].data, f"graph1.param.{k}", tol=1e-5,
                     outlier_frac=2e-3)
    assert_close(graphed.xys_grad_norm, eager.xys_grad_norm, "graph1.xys_grad_norm", tol=1e-5, outlier_frac=1e-3)


def test_overflowed_step_is_a_noop_and_recovers():
    from fusionsense_b200.graph_step import GraphedDNSplatterStep

    eager, graphed, targets = _pair()
    runner = GraphedDNSplatterStep(graphed, targets, capacity=1000)  # far too small
    before = {k: v.data.clone() for k, v in graphed.gauss_params.items()}
    runner.train_iteration(0)
    info = runner.poll()
    assert info["new_overflows"] == 1 and graphed.step == 3000
    for k, v in graphed.gauss_params.items():
        assert torch.equal(v.data, before[k]), k
    p = graphed.optimizers["means"].param_groups[0]["params"][0]
    assert float(graphed.optimizers["means"].state[p]["step"]) == 0.0
    assert float(graphed.optimizers["means"].state[p]["exp_avg"].abs().max()) == 0.0
    assert runner.capacity > info["n_isects_normals"]
    # the re-captured step now applies and matches the eager one
    le = float(eager.train_iteration(0, targets[0]))
    runner.train_iteration(0)
    info = runner.poll()
    assert info["new_overflows"] == 0 and runner.captures == 2 and graphed.step == 3001
    assert info["loss"] == pytest.approx(le, rel=1e-5)


def test_staged_targets_feed_the_replay():
    from fusionsense_b200.graph_step import GraphedDNSplatterStep

    eager, graphed, targets = _pair()
    runner = GraphedDNSplatterStep(graphed, targets)
    other = {k: (t * 0.5).cpu().pin_memory() for k, t in targets[2].items()}
    nbytes = runner.stage(2, other)
    assert nbytes == sum(t.numel() * 4 for t in other.values())
    le = float(eager.train_iteration(2, {k: t.cuda() for k, t in other.items()}))
    runner.train_iteration(2)
    assert runner.poll()["loss"] == pytest.approx(le, rel=1e-5)


def test_prefetched_targets_and_async_result_ring_match_the_blocking_path():
    """bench.py's e2e leg: step i + 1's targets are copied on the copy stream while step i runs, and step i's result
    is read one step late from a pinned ring.  Same losses as staging + polling synchronously."""
    from fusionsense_b200.graph_step import GraphedDNSplatterStep

    a, b, targets = _pair()
    ra, rb = GraphedDNSplatterStep(a, targets), GraphedDNSplatterStep(b, targets)
    # every step trains on a DIFFERENT host-side version of its view, so a late or early copy changes the loss
    order = [0, 1, 2, 0, 1, 2, 1]
    host = [{k: (t * (1.0 - 0.05 * i)).cpu().pin_memory() for k, t in targets[v].items()} for i, v in enumerate(order)]
    blocking = []
    for i, v in enumerate(order):
        ra.stage(v, host[i])
        ra.train_iteration(v)
        blocking.append(ra.poll()["loss"])
    piped = []
    rb.stage_async(order[0], host[0])
    for i, v in enumerate(order):
        rb.train_iteration(v)
        if i + 1 < len(order):
            rb.stage_async(order[i + 1], host[i + 1])
        prev = rb.read_result_async()
        if prev is not None:
            piped.append(prev[0])
    piped.append(rb.poll()["loss"])
    assert len(piped) == len(order)
    for x, y in zip(piped, blocking):
        assert x == pytest.approx(y, rel=2e-4)


@pytest.mark.parametrize("captured", [True, False])
def test_step_with_grad_sync_matches_single_gpu_step(captured):
    """The N > 1 structures on one GPU, where the exchange is just the pack into the flat buffer: the exchange captured
    inside the one graph (default), and main graph -> eager exchange -> tail graph.  Same losses and parameters as
    the single-GPU step, overflow still a no-op."""
    from fusionsense_b200.dist import GradSync
    from fusionsense_b200.graph_step import GraphedDNSplatterStep

    single, split, targets = _pair()
    r1 = GraphedDNSplatterStep(single, targets)
    r2 = GraphedDNSplatterStep(split, targets, grad_sync=GradSync())
    r2.capture_collective = captured
    for v in [0, 2, 1, 1, 0]:
        r1.train_iteration(v)
        r2.train_iteration(v)
        assert r2.poll()["loss"] == pytest.approx(r1.poll()["loss"], rel=2e-4)
    assert (r2.graph_tail is None) == captured and r1.graph_tail is None and r2.captures == 1
    assert r2.poll()["overflowed_steps"] == 0 and split.step == single.step
    for k in single.gauss_params:
        assert_close(split.gauss_params[k].data, single.gauss_params[k].data, f"split.param.{k}", tol=2e-3,
                     outlier_frac=2e-2)
    # overflow in the split structure: the tail graph sees the flag the exchange step produced and skips Adam
    _, small, targets = _pair()
    r3 = GraphedDNSplatterStep(small, targets, capacity=1000, grad_sync=GradSync())
    r3.capture_collective = captured
    before = {k: v.data.clone() for k, v in small.gauss_params.items()}
    r3.train_iteration(0)
    assert r3.poll()["new_overflows"] == 1
    for k, v in small.gauss_params.items():
        assert torch.equal(v.data, before[k]), k


def test_graph_step_survives_refine_steps_that_drop_the_statistics():
    """refinement_after ends every refine step with xys_grad_norm = vis_counts = max_2Dsize = None, also when it
    neither densifies nor culls (step % 3000 in {0, 100} with the defaults): the parameters keep their identity
    there, so only the statistics tell the captured step that it must be re-captured (round-1 advisor finding)."""
    from fusionsense_b200.graph_step import GraphedDNSplatterStep

    _, graphed, targets = _pair()
    graphed.step = 2998
    runner = GraphedDNSplatterStep(graphed, targets)
    n0 = graphed.num_points
    for i in range(2):
        runner.train_iteration(i % 3)
    assert graphed.step == 3000 and runner.captures == 1
    runner.poll()
    deleted = graphed.refinement_after()  # step 3000: 3000 % 3000 = 0 -> no densification, no cull, stats dropped
    assert deleted is None and graphed.num_points == n0 and graphed.xys_grad_norm is None
    for i in range(3):
        runner.train_iteration(i % 3)
    assert runner.captures == 2, "dropping the statistics must re-capture"
    assert graphed.xys_grad_norm is not None and float(graphed.xys_grad_norm.abs().sum()) > 0
    assert float(graphed.vis_counts.max()) == 4.0  # ones + three accumulated iterations, not stale memory
    # and a densifying refine step right after works on live statistics
    graphed.step = 3200
    runner.poll()
    graphed.refinement_after()
    runner.train_iteration(0)
    assert runner.captures == 3 and runner.poll()["overflowed_steps"] == 0


def test_camera_batch_per_iteration_accumulates_like_separate_backwards():
    """views_per_iter = 3: one captured iteration renders three views, accumulates their gradients (each loss / 3),
    counts all three in the densification statistics and steps Adam once — compared with the eager model doing the
    same by hand."""
    from fusionsense_b200.graph_step import GraphedDNSplatterStep
    from fusionsense_b200.optim import fused_step

    eager, graphed, targets = _pair()
    runner = GraphedDNSplatterStep(graphed, targets, views_per_iter=3)
    for views in ([0, 1, 2], [2, 0, 1]):
        for opt in eager.optimizers.values():
            opt.zero_grad(set_to_none=True)
        losses = []
        for v in views:
            out = eager.get_outputs(v)
            loss = eager.get_loss_dict(out, targets[v])["main_loss"]
            (loss / 3).backward()
            eager.after_train()
            losses.append(float(loss))
        eager.optimizers["means"].param_groups[0]["lr"] = eager._means_lr()
        fused_step(eager.optimizers.values())
        eager.step += 1
        runner.train_iteration(views)
        info = runner.poll()
        assert info["loss"] == pytest.approx(sum(losses) / 3, rel=2e-5) and info["overflowed_steps"] == 0
    assert runner.captures == 1 and graphed.step == eager.step
    for k in eager.gauss_params:
        assert_close(graphed.gauss_params[k].data, eager.gauss_params[k].data, f"graph3.param.{k}", tol=1e-5,
                     outlier_frac=1e-4)
    assert torch.equal(graphed.vis_counts, eager.vis_counts)
    assert_close(graphed.xys_grad_norm, eager.xys_grad_norm, "graph3.xys_grad_norm", tol=1e-5, outlier_frac=1e-4)


@pytest.mark.parametrize("n", [0, 1, 15, 16, 4099, 3 * 640 * 480 + 5])
def test_u8_targets_convert_to_torch_bits(n):
    """fsb_u8_to_unit_float == `image.float() / 255.0` on the device (splatfacto get_gt_img) and numpy's
    `x.astype("float32") / 255.0` (dn_dataset.py:205), bit for bit, on aligned and unaligned starts and ragged tails."""
    from fusionsense_b200.compose import u8_to_unit_float

    g = torch.Generator().manual_seed(n)
    base = torch.randint(0, 256, (n + 3,), dtype=torch.uint8, generator=g).to(DEV)
    for off in (0, 1, 3):
        src = base[off:off + n]
        if n > 256:
            src[:256] = torch.arange(256, dtype=torch.uint8, device=DEV)  # every byte value at least once
        src = src.contiguous() if off == 0 else src  # a slice of a 1-D tensor is contiguous but unaligned
        # on the device torch divides by a Python scalar through the fp32 reciprocal (get_gt_img runs there) ...
        assert torch.equal(u8_to_unit_float(src), src.float() / 255.0)
        # ... numpy divides (dn_dataset.py:205)
        want = torch.from_numpy(src.cpu().numpy().astype("float32") / 255.0).to(DEV)
        assert torch.equal(u8_to_unit_float(src, recip=False), want)


def test_staged_u8_targets_equal_resident_float_targets():
    """8-bit host targets staged through stage_async / stage land in the resident slots as the float32 values the
    reference computes after its own copy; a replay on them gives the loss of the float32 targets."""
    from fusionsense_b200.graph_step import GraphedDNSplatterStep

    from fusionsense_b200.graph_step import eight_bit_targets

    eager, graphed, raw = _pair()
    both = {v: eight_bit_targets(raw[v]) for v in range(3)}
    targets = {v: both[v][0] for v in range(3)}
    host = {v: both[v][1] for v in range(3)}
    assert host[0]["image"].dtype == torch.uint8 and host[0]["normal"].dtype == torch.uint8
    assert host[0]["sensor_depth"].dtype == torch.float32 and host[0]["image"].is_pinned()
    # the resident float32 targets are what the reference's loaders make of the bytes
    assert torch.equal(targets[1]["image"], host[1]["image"].to(DEV).float() / 255.0)
    assert torch.equal(targets[1]["normal"].cpu(), torch.from_numpy(host[1]["normal"].numpy().astype("float32") / 255.0))
    runner = GraphedDNSplatterStep(graphed, targets)
    want = {k: t.clone() for k, t in runner.targets.items()}
    for t in runner.targets.values():
        t.zero_()
    moved = runner.stage_async(0, host[0]) + runner.stage(1, host[1])
    runner.stage_async(2, host[2])
    px = 256 * 192
    assert moved == 2 * px * (3 + 3 + 4)
    runner.train_iteration(0)
    runner.train_iteration(2)
    loss = runner.poll()["loss"]
    torch.cuda.synchronize()
    for k in want:
        assert torch.equal(runner.targets[k], want[k]), k
    eager.train_iteration(0, targets[0])
    assert loss == pytest.approx(float(eager.train_iteration(2, targets[2])), rel=2e-4)
