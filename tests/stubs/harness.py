"""Drive the reference's unmodified DNSplatterModel (baseline/_ref) the way nerfstudio's Trainer does, on top of the
stub nerfstudio and this repository's gsplat drop-in.  TEST / BENCH INFRASTRUCTURE (tests/stubs/__init__.py)."""
from __future__ import annotations

import torch

from . import install


def fusionsense_model_config(ref_model, step_config=None):
    """configs/config.py:3-39 + scripts/train.py:117-145 over DNSplatterModelConfig: the values DNSplatterStepConfig
    names "effective FusionSense values"."""
    from fusionsense_b200.dn_step import DNSplatterStepConfig

    s = step_config or DNSplatterStepConfig()
    return ref_model.DNSplatterModelConfig(
        use_depth_loss=s.use_depth_loss, depth_loss_type=ref_model.DepthLossType.EdgeAwareLogL1,
        depth_tolerance=s.depth_tolerance, sensor_depth_lambda=s.sensor_depth_lambda,
        use_depth_smooth_loss=s.use_depth_smooth_loss, smooth_loss_lambda=s.smooth_loss_lambda,
        use_normal_loss=s.use_normal_loss, use_normal_tv_loss=s.use_normal_tv_loss, normal_lambda=s.normal_lambda,
        normal_supervision="mono", two_d_gaussians=s.two_d_gaussians, use_binary_opacities=s.use_binary_opacities,
        binary_opacities_threshold=s.binary_opacities_threshold, warmup_length=s.warmup_length,
        stop_split_at=s.stop_split_at, sh_degree=s.sh_degree, ssim_lambda=s.ssim_lambda)


def gl_scene(scene):
    """The scene with the camera matrices a nerfstudio `Cameras` holds (OpenGL c2w) and the view matrices
    `get_viewmat` derives from them, so the reference model and DNSplatterStep see bit-identical cameras."""
    install()
    from nerfstudio.models.splatfacto import get_viewmat

    from fusionsense_b200.synthetic import Scene

    c2w_gl = scene.c2w.clone()
    c2w_gl[:, :3, 1:3] *= -1.0  # OpenCV -> OpenGL camera axes
    viewmats = get_viewmat(c2w_gl[:, :3, :4])
    return Scene(scene.means, scene.scales, scene.quats, scene.opacities, scene.features_dc, scene.features_rest,
                 viewmats, c2w_gl, scene.Ks, scene.width, scene.height)


def build_reference_model(scene, step: int, device="cuda", step_config=None):
    """-> (dn_splatter.dn_model module, DNSplatterModel in train mode holding the scene's parameters)."""
    install()
    import dn_splatter.dn_model as ref_model  # the reference file, executed as it is

    N = scene.N
    torch.manual_seed(0)
    # populate_modules runs a CPU k-NN over the seed points for the initial scales: a small seed cloud is enough,
    # every parameter is replaced by the scene's right after
    n_seed = min(N, 2048)
    seed = (scene.means[:n_seed].clone(), torch.full((n_seed, 3), 128.0))
    model = ref_model.DNSplatterModel(fusionsense_model_config(ref_model, step_config),
                                      num_train_data=scene.viewmats.shape[0], seed_points=seed).to(device)
    for name in ("means", "scales", "quats", "features_dc", "features_rest", "opacities"):
        model.gauss_params[name] = torch.nn.Parameter(getattr(scene, name).clone().to(device))
    model.gauss_params["normals"] = torch.nn.Parameter(torch.zeros(N, 3, device=device))
    model.train()
    model.step = step
    return ref_model, model


def camera_for(scene, cam_idx: int, device="cuda"):
    from nerfstudio.cameras.cameras import Cameras

    K = scene.Ks[cam_idx]
    return Cameras(scene.c2w[cam_idx:cam_idx + 1, :3, :4].to(device), K[0, 0].item(), K[1, 1].item(), K[0, 2].item(),
                   K[1, 2].item(), scene.width, scene.height, metadata={"cam_idx": cam_idx})


def build_optimizers(model, lrs, fused: bool = False):
    """One torch.optim.Adam(lr, eps=1e-15) per Gaussian parameter group (dn_config.py:36-75) behind nerfstudio's
    `Optimizers` (the `normals` group exists and never receives a gradient, dn_config.py:69-74).
    `fused`: the same method config passed through fusionsense_b200.optim.use_fused_adam."""
    from nerfstudio.engine.optimizers import AdamOptimizerConfig, Optimizers

    groups = model.get_gaussian_param_groups()
    config = {name: {"optimizer": AdamOptimizerConfig(lr=lrs.get(name, 1e-3), eps=1e-15), "scheduler": None}
              for name in groups}
    if fused:
        from fusionsense_b200.optim import use_fused_adam

        use_fused_adam(config)
    return Optimizers(config, groups)


def train_iteration(model, optimizers, camera, batch, step: int):
    """Trainer.train_iteration: step callback, zero_grad, forward + losses, backward, optimizer
    steps, after_train."""
    model.step_cb(step)
    optimizers.zero_grad_all()
    outputs = model.get_outputs(camera)
    loss_dict = model.get_loss_dict(outputs, dict(batch))
    loss = sum(loss_dict.values())
    loss.backward()
    optimizers.optimizer_step_all()
    model.after_train(step)
    return loss
