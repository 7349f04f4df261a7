"""The reference's OWN `dn_splatter/dn_model.py`, unmodified, running on this repository's `gsplat` drop-in.

`DNSplatterModel` (baseline/_ref/dn_splatter/dn_model.py = /root/reference/dn_splatter/dn_model.py as installed by the
sanctioned `pip install --no-deps --target baseline/_ref`) is imported against stub `nerfstudio` / `torchmetrics`
packages (tests/stubs, restated from SURVEY.md Appendix A.7: neither is installed in this image), instantiated on a
synthetic scene and driven through `get_outputs` (dn_model.py:469-671: `gsplat.rendering.rasterization` at :570-591 and
`gsplat.rasterize_gaussians` at :644-653 resolve to libfsb200.so), `get_loss_dict` (:673-925) and backward.  Its
outputs, loss and parameter gradients are compared with `fusionsense_b200.dn_step.DNSplatterStep` — the literal
restatement (fused_* off) and the default fused step — and with the CPU oracle.

Skipped when baseline/_ref does not hold the reference package (it is git-ignored; gpurun ships it to the GPU box)."""
import sys

import pytest
import torch

from fusionsense_b200.synthetic import Scene, make_scene
from oracle import dn_losses_ref as torch_losses
from tests import stubs
from tests.parity import assert_close

needs_ref = pytest.mark.skipif(not stubs.reference_available(), reason="baseline/_ref/dn_splatter not installed")


def _import_reference():
    stubs.install()
    import dn_splatter.dn_model as ref_model  # the reference file, executed as it is

    assert "baseline/_ref" in ref_model.__file__.replace("\\", "/")
    return ref_model


@needs_ref
def test_reference_module_imports_against_stubs_and_shim():
    ref_model = _import_reference()
    import gsplat

    assert gsplat.rasterization.__module__.startswith("fusionsense_b200")
    assert ref_model.rasterization is gsplat.rendering.rasterization
    assert ref_model.rasterize_gaussians is gsplat.rasterize_gaussians
    cfg = ref_model.DNSplatterModelConfig()
    assert cfg.background_color == "white" and cfg.warmup_length == 500
    assert sys.modules["nerfstudio"].__version__.endswith("stub")


def _gl_scene(n=20000, W=320, H=240, views=3):
    from tests.stubs.harness import gl_scene

    return gl_scene(make_scene(n, W, H, n_views=views, cfg_id=91, kind="bunny", fx=300.0))


@needs_ref
@pytest.mark.gpu
def test_reference_dn_model_runs_on_the_shim_and_matches_dn_step():
    _import_reference()
    from fusionsense_b200.dn_step import DNSplatterStep, DNSplatterStepConfig

    dev = "cuda"
    scene = _gl_scene()
    step = 3001  # past warm-up, SH degree 3, binary opacities on; not a multiple of 100 (dn_model.py:905 writes a jpg)

    # ---- the reference model -----------------------------------------------------------------------------------
    from tests.stubs.harness import build_reference_model, camera_for

    _, model = build_reference_model(scene, step, dev)
    cam_idx = 1
    camera = camera_for(scene, cam_idx, dev)

    # ---- DNSplatterStep twins on the same parameters ----------------------------------------------------------------
    lit = DNSplatterStep(scene, DNSplatterStepConfig(fused_optimizer=False, fused_losses=False, fused_glue=False,
                                                     fused_outputs=False), device=dev, step=step,
                         torch_losses=torch_losses)
    fused = DNSplatterStep(scene, DNSplatterStepConfig(), device=dev, step=step)
    batch = lit.render_targets(0)  # some other view's render as this view's ground truth: non-trivial losses

    launches0 = _launches()
    out_ref = model.get_outputs(camera)
    assert _launches() > launches0, "the reference's gsplat calls must reach libfsb200.so"
    loss_ref = sum(model.get_loss_dict(out_ref, {k: v.clone() for k, v in batch.items()}).values())
    loss_ref.backward()
    # after_train (splatfacto, inherited): needs meta["means2d"].absgrad from our backward
    model.after_train(step)
    assert model.xys_grad_norm is not None and float(model.xys_grad_norm.sum()) > 0

    for tag, twin, tol_img, tol_grad in (("literal", lit, 1e-6, 2e-5), ("fused", fused, 1e-5, 1e-4)):
        out = twin.get_outputs(cam_idx)
        ld = twin.get_loss_dict(out, batch)
        loss = ld["main_loss"] + ld["scale_reg"]
        loss.backward()
        for k in ("rgb", "depth", "normal", "accumulation"):
            assert_close(out[k], out_ref[k], f"ref_dn_model.{tag}.{k}", tol=tol_img, outlier_frac=1e-4)
        assert float(loss) == pytest.approx(float(loss_ref), rel=2e-5), tag
        for name in ("means", "scales", "quats", "features_dc", "features_rest", "opacities"):
            assert_close(twin.gauss_params[name].grad, model.gauss_params[name].grad, f"ref_dn_model.{tag}.grad.{name}",
                         tol=tol_grad, outlier_frac=2e-3)
    # densification statistics: the reference's after_train vs the step's fused kernel
    fused.after_train()
    assert_close(fused.xys_grad_norm, model.xys_grad_norm, "ref_dn_model.xys_grad_norm", tol=1e-4, outlier_frac=2e-3)
    assert torch.equal(fused.vis_counts, model.vis_counts)
    assert_close(fused.max_2Dsize, model.max_2Dsize, "ref_dn_model.max_2Dsize", tol=1e-6)


@needs_ref
@pytest.mark.gpu
def test_reference_dn_model_against_cpu_oracle():
    """The same reference file, once on libfsb200.so (GPU) and once on the CPU oracle standing in for gsplat."""
    ref_model = _import_reference()
    from oracle import gsplat_ref as oracle

    from fusionsense_b200.dn_step import DNSplatterStep, DNSplatterStepConfig

    scene = _gl_scene(n=6000, W=256, H=192)
    step = 3001
    gpu = DNSplatterStep(scene, DNSplatterStepConfig(), device="cuda", step=step)
    batch = gpu.render_targets(0)
    cpu = DNSplatterStep(scene, DNSplatterStepConfig(fused_optimizer=False, stop_split_at=0), device="cpu", step=step,
                         gsplat_module=oracle, torch_losses=torch_losses)
    out_c = cpu.get_outputs(1)
    loss_c = cpu.get_loss_dict(out_c, {k: v.cpu() for k, v in batch.items()})["main_loss"]
    loss_c.backward()

    from tests.stubs.harness import build_reference_model, camera_for

    _, model = build_reference_model(scene, step, "cuda")
    camera = camera_for(scene, 1, "cuda")
    out_ref = model.get_outputs(camera)
    loss_ref = sum(model.get_loss_dict(out_ref, {k: v.clone() for k, v in batch.items()}).values())
    loss_ref.backward()
    for k in ("rgb", "depth", "normal", "accumulation"):
        assert_close(out_ref[k], out_c[k], f"ref_dn_model.vs_oracle.{k}", tol=1e-4, outlier_frac=1e-3)
    assert float(loss_ref) == pytest.approx(float(loss_c), rel=1e-4)
    for name in ("means", "scales", "quats", "opacities", "features_dc"):
        a, b = model.gauss_params[name].grad.cpu(), cpu.gauss_params[name].grad
        assert float((a - b).norm() / (b.norm() + 1e-20)) < 5e-3, name


def _launches():
    from fusionsense_b200._abi import lib

    return lib.fsb_launch_count()


@needs_ref
@pytest.mark.gpu
@pytest.mark.parametrize("fused_adam", [False, True])
def test_reference_train_iterations_track_the_fused_step(fused_adam):
    """Three Trainer.train_iteration()s of the reference model (stub `Optimizers`: one torch.optim.Adam(eps=1e-15) per
    group, dn_config.py:36-75) against three iterations of the fused step (one fused Adam launch): losses and
    parameters stay together.  `fused_adam`: the method config's optimizers swapped for FusedAdamOptimizerConfig
    (optim.use_fused_adam), i.e. the reference model trained by our optimizer behind nerfstudio's `Optimizers`."""
    _import_reference()
    from fusionsense_b200.dn_step import DNSplatterStep, DNSplatterStepConfig
    from tests.stubs.harness import build_optimizers, build_reference_model, camera_for, train_iteration

    dev = "cuda"
    scene = _gl_scene(n=8000, W=256, H=192)
    step0 = 3001
    fused = DNSplatterStep(scene, DNSplatterStepConfig(), device=dev, step=step0)
    targets = {v: fused.render_targets(v) for v in range(3)}
    _, model = build_reference_model(scene, step0, dev)
    lrs = dict(fused.config.lrs)
    opts = build_optimizers(model, lrs, fused=fused_adam)
    assert type(opts.optimizers["means"]).__name__ == ("FusedAdam" if fused_adam else "Adam")
    for i, v in enumerate([0, 2, 1]):
        opts.optimizers["means"].param_groups[0]["lr"] = fused._means_lr()  # ExponentialDecayScheduler on means
        loss_ref = train_iteration(model, opts, camera_for(scene, v, dev), targets[v], step0 + i)
        loss = fused.train_iteration(v, targets[v])
        assert float(loss) == pytest.approx(float(loss_ref), rel=1e-4), i
    for name in ("means", "scales", "quats", "features_dc", "features_rest", "opacities"):
        assert_close(fused.gauss_params[name], model.gauss_params[name], f"ref_dn_model.train3.{name}", tol=1e-4,
                     outlier_frac=2e-3)
    assert_close(fused.xys_grad_norm, model.xys_grad_norm, "ref_dn_model.train3.xys_grad_norm", tol=1e-3, outlier_frac=5e-3)


@needs_ref
@pytest.mark.gpu
def test_reference_level_surface_search_with_the_gpu_knn_drop_in():
    """f2: the reference's own `compute_level_surface_points` (dn_model.py:1705-1946: render,
    back-project, `knn_sk(self.means, points, 16)`, 21 density samples per ray, level crossings) run twice on the same
    model and camera — with the reference's sklearn `knn_sk` and with `fusionsense_b200.knn.knn_sk` bound in its place
    (the one-line patch of INTEGRATION.md).  Exact neighbours -> every output tensor is bit-identical."""
    import random

    ref_model = _import_reference()
    from fusionsense_b200 import knn as fsb_knn
    from tests.stubs.harness import build_reference_model, camera_for

    dev = "cuda"
    scene = _gl_scene(n=20000, W=160, H=120)
    _, model = build_reference_model(scene, 3001, dev)
    g = torch.Generator().manual_seed(3)
    model.gauss_params["normals"] = torch.nn.Parameter(
        torch.nn.functional.normalize(torch.randn(scene.N, 3, generator=g), dim=-1).to(dev))
    camera = camera_for(scene, 1, dev)
    sklearn_knn = ref_model.knn_sk
    assert sklearn_knn.__module__ == "dn_splatter.utils.knn"
    calls = []

    def ours(x, y, k):
        calls.append((tuple(x.shape), tuple(y.shape), k))
        return fsb_knn.knn_sk(x, y, k)

    random.seed(11)
    want = model.compute_level_surface_points(camera, num_samples=4000)
    try:
        ref_model.knn_sk = ours
        random.seed(11)
        got = model.compute_level_surface_points(camera, num_samples=4000)
    finally:
        ref_model.knn_sk = sklearn_knn
    assert len(calls) == 1 and calls[0][0] == (scene.N, 3) and calls[0][2] == 16 and calls[0][1][0] > 1000
    assert set(got) == set(want) == {0.1, 0.3, 0.5}
    n_pts = 0
    for level in want:
        for key in ("points", "normals", "colors"):
            assert torch.equal(got[level][key], want[level][key]), (level, key)
        n_pts += want[level]["points"].shape[0]
    assert n_pts > 1000  # the search found surfaces: the comparison is not vacuous
