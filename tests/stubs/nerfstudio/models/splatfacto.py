"""Stub of nerfstudio 1.1.3 `nerfstudio.models.splatfacto` restated from SURVEY.md Appendix A.7: the parts
`DNSplatterModel` (/root/reference/dn_splatter/dn_model.py) inherits.  TEST INFRASTRUCTURE (tests/stubs/__init__.py)."""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Dict, List, Literal, Optional, Type

import numpy as np
import torch
from torch import nn

from nerfstudio.cameras.camera_optimizers import CameraOptimizerConfig

SH_C0 = 0.28209479177387814


def RGB2SH(rgb):
    return (rgb - 0.5) / SH_C0


def SH2RGB(sh):
    return sh * SH_C0 + 0.5


def get_viewmat(optimized_camera_to_world):
    """c2w [B,3,4] (OpenGL) -> world-to-camera [B,4,4] (OpenCV): R <- R diag(1,-1,-1); viewmat = [[R^T, -R^T t],[0 0 0 1]]."""
    R = optimized_camera_to_world[:, :3, :3]
    T = optimized_camera_to_world[:, :3, 3:4]
    R = R * torch.tensor([[[1, -1, -1]]], device=R.device, dtype=R.dtype)
    R_inv = R.transpose(1, 2)
    T_inv = -torch.bmm(R_inv, T)
    viewmat = torch.zeros(R.shape[0], 4, 4, device=R.device, dtype=R.dtype)
    viewmat[:, 3, 3] = 1.0
    viewmat[:, :3, :3] = R_inv
    viewmat[:, :3, 3:4] = T_inv
    return viewmat


def random_quat_tensor(N):
    u, v, w = torch.rand(N), torch.rand(N), torch.rand(N)
    return torch.stack([torch.sqrt(1 - u) * torch.sin(2 * np.pi * v), torch.sqrt(1 - u) * torch.cos(2 * np.pi * v),
                        torch.sqrt(u) * torch.sin(2 * np.pi * w), torch.sqrt(u) * torch.cos(2 * np.pi * w)], dim=-1)


@dataclass
class SplatfactoModelConfig:
    _target: Type = field(default_factory=lambda: SplatfactoModel)
    warmup_length: int = 500
    refine_every: int = 100
    resolution_schedule: int = 3000
    background_color: Literal["random", "black", "white"] = "random"
    num_downscales: int = 2
    cull_alpha_thresh: float = 0.1
    cull_scale_thresh: float = 0.5
    continue_cull_post_densification: bool = True
    reset_alpha_every: int = 30
    densify_grad_thresh: float = 0.0008
    densify_size_thresh: float = 0.01
    n_split_samples: int = 2
    sh_degree_interval: int = 1000
    cull_screen_size: float = 0.15
    split_screen_size: float = 0.05
    stop_screen_size_at: int = 4000
    random_init: bool = False
    num_random: int = 50000
    random_scale: float = 10.0
    ssim_lambda: float = 0.2
    stop_split_at: int = 15000
    sh_degree: int = 3
    use_scale_regularization: bool = False
    max_gauss_ratio: float = 10.0
    output_depth_during_training: bool = False
    rasterize_mode: Literal["classic", "antialiased"] = "classic"
    camera_optimizer: CameraOptimizerConfig = field(default_factory=lambda: CameraOptimizerConfig(mode="off"))

    def setup(self, **kwargs):
        return self._target(self, **kwargs)


class SplatfactoModel(nn.Module):
    config: SplatfactoModelConfig

    def __init__(self, config, scene_box=None, num_train_data: int = 1, seed_points=None, **kwargs):
        super().__init__()
        self.config = config
        self.scene_box = scene_box
        self.num_train_data = num_train_data
        self.seed_points = seed_points
        self.kwargs = kwargs
        self.device_indicator_param = nn.Parameter(torch.empty(0))
        self.populate_modules()

    # ---- Model plumbing ---------------------------------------------------------------------------------------
    @property
    def device(self):
        return self.device_indicator_param.device

    def forward(self, camera):
        return self.get_outputs(camera)

    def step_cb(self, step):
        self.step = step

    @property
    def num_points(self):
        return self.means.shape[0]

    @property
    def means(self):
        return self.gauss_params["means"]

    @property
    def scales(self):
        return self.gauss_params["scales"]

    @property
    def quats(self):
        return self.gauss_params["quats"]

    @property
    def features_dc(self):
        return self.gauss_params["features_dc"]

    @property
    def features_rest(self):
        return self.gauss_params["features_rest"]

    @property
    def opacities(self):
        return self.gauss_params["opacities"]

    @property
    def colors(self):
        if self.config.sh_degree > 0:
            return SH2RGB(self.features_dc)
        return torch.sigmoid(self.features_dc)

    def k_nearest_sklearn(self, x: torch.Tensor, k: int):
        from sklearn.neighbors import NearestNeighbors

        x_np = x.cpu().numpy()
        nn_model = NearestNeighbors(n_neighbors=k + 1, algorithm="auto", metric="euclidean").fit(x_np)
        distances, indices = nn_model.kneighbors(x_np)
        return distances[:, 1:].astype(np.float32), indices[:, 1:].astype(np.float32)

    def _get_downscale_factor(self):
        if self.training:
            return 2 ** max((self.config.num_downscales - self.step // self.config.resolution_schedule), 0)
        return 1

    def _downscale_if_required(self, image):
        d = self._get_downscale_factor()
        if d > 1:
            import torch.nn.functional as F

            return F.interpolate(image.permute(2, 0, 1)[None], scale_factor=1.0 / d, mode="bilinear")[0].permute(1, 2, 0)
        return image

    def _get_background_color(self):
        if self.config.background_color == "random":
            background = torch.rand(3, device=self.device) if self.training else self.background_color.to(self.device)
        elif self.config.background_color == "white":
            background = torch.ones(3, device=self.device)
        elif self.config.background_color == "black":
            background = torch.zeros(3, device=self.device)
        else:
            raise ValueError(self.config.background_color)
        return background

    @staticmethod
    def get_empty_outputs(width: int, height: int, background: torch.Tensor):
        rgb = background.repeat(height, width, 1)
        depth = background.new_ones(*rgb.shape[:2], 1) * 10
        accumulation = background.new_zeros(*rgb.shape[:2], 1)
        return {"rgb": rgb, "depth": depth, "accumulation": accumulation, "background": background}

    def get_gt_img(self, image: torch.Tensor):
        if image.dtype == torch.uint8:
            image = image.float() / 255.0
        gt_img = self._downscale_if_required(image)
        return gt_img.to(self.device)

    def composite_with_background(self, image, background):
        if image.shape[2] == 4:
            alpha = image[..., -1].unsqueeze(-1).repeat((1, 1, 3))
            return alpha * image[..., :3] + (1 - alpha) * background
        return image

    def get_gaussian_param_groups(self) -> Dict[str, List[nn.Parameter]]:
        return {name: [self.gauss_params[name]]
                for name in ["means", "scales", "quats", "features_dc", "features_rest", "opacities"]}

    def get_param_groups(self):
        gps = self.get_gaussian_param_groups()
        self.camera_optimizer.get_param_groups(param_groups=gps)
        return gps

    # ---- losses -----------------------------------------------------------------------------------------------
    def get_metrics_dict(self, outputs, batch):
        gt_rgb = self.composite_with_background(self.get_gt_img(batch["image"]), outputs["background"])
        metrics_dict = {"psnr": self.psnr(outputs["rgb"], gt_rgb), "gaussian_count": self.num_points}
        self.camera_optimizer.get_metrics_dict(metrics_dict)
        return metrics_dict

    def get_loss_dict(self, outputs, batch, metrics_dict=None):
        gt_img = self.composite_with_background(self.get_gt_img(batch["image"]), outputs["background"])
        pred_img = outputs["rgb"]
        if "mask" in batch:
            mask = self._downscale_if_required(batch["mask"]).to(self.device)
            assert mask.shape[:2] == gt_img.shape[:2] == pred_img.shape[:2]
            gt_img = gt_img * mask
            pred_img = pred_img * mask
        Ll1 = torch.abs(gt_img - pred_img).mean()
        simloss = 1 - self.ssim(gt_img.permute(2, 0, 1)[None, ...], pred_img.permute(2, 0, 1)[None, ...])
        if self.config.use_scale_regularization and self.step % 10 == 0:
            scale_exp = torch.exp(self.scales)
            scale_reg = (torch.maximum(scale_exp.amax(dim=-1) / scale_exp.amin(dim=-1),
                                       torch.tensor(self.config.max_gauss_ratio)) - self.config.max_gauss_ratio)
            scale_reg = 0.1 * scale_reg.mean()
        else:
            scale_reg = torch.tensor(0.0).to(self.device)
        loss_dict = {"main_loss": (1 - self.config.ssim_lambda) * Ll1 + self.config.ssim_lambda * simloss,
                     "scale_reg": scale_reg}
        if self.training:
            self.camera_optimizer.get_loss_dict(loss_dict)
        return loss_dict

    # ---- densification bookkeeping -------------------------------------------------------------------------------
    def after_train(self, step: int):
        assert step == self.step
        if self.step >= self.config.stop_split_at:
            return
        with torch.no_grad():
            visible_mask = (self.radii > 0).flatten()
            grads = self.xys.absgrad[0][visible_mask].norm(dim=-1)
            if self.xys_grad_norm is None:
                self.xys_grad_norm = torch.zeros(self.num_points, device=self.device, dtype=torch.float32)
                self.vis_counts = torch.ones(self.num_points, device=self.device, dtype=torch.float32)
            assert self.vis_counts is not None
            self.vis_counts[visible_mask] += 1
            self.xys_grad_norm[visible_mask] += grads
            if self.max_2Dsize is None:
                self.max_2Dsize = torch.zeros_like(self.radii, dtype=torch.float32)
            newradii = self.radii.detach()[visible_mask]
            self.max_2Dsize[visible_mask] = torch.maximum(
                self.max_2Dsize[visible_mask], newradii / float(max(self.last_size[0], self.last_size[1])))

    def remove_from_optim(self, optimizer, deleted_mask, new_params):
        assert len(new_params) == 1
        param = optimizer.param_groups[0]["params"][0]
        param_state = optimizer.state[param]
        del optimizer.state[param]
        if "exp_avg" in param_state:
            param_state["exp_avg"] = param_state["exp_avg"][~deleted_mask]
            param_state["exp_avg_sq"] = param_state["exp_avg_sq"][~deleted_mask]
        del optimizer.param_groups[0]["params"][0]
        del optimizer.param_groups[0]["params"]
        optimizer.param_groups[0]["params"] = new_params
        optimizer.state[new_params[0]] = param_state

    def dup_in_optim(self, optimizer, dup_mask, new_params, n=2):
        param = optimizer.param_groups[0]["params"][0]
        param_state = optimizer.state[param]
        if "exp_avg" in param_state:
            repeat_dims = (n,) + tuple(1 for _ in range(param_state["exp_avg"].dim() - 1))
            param_state["exp_avg"] = torch.cat(
                [param_state["exp_avg"], torch.zeros_like(param_state["exp_avg"][dup_mask.squeeze()]).repeat(*repeat_dims)], dim=0)
            param_state["exp_avg_sq"] = torch.cat(
                [param_state["exp_avg_sq"], torch.zeros_like(param_state["exp_avg_sq"][dup_mask.squeeze()]).repeat(*repeat_dims)], dim=0)
        del optimizer.state[param]
        optimizer.state[new_params[0]] = param_state
        optimizer.param_groups[0]["params"] = new_params
        del param

    def cull_gaussians(self, extra_cull_mask: Optional[torch.Tensor] = None):
        culls = (torch.sigmoid(self.opacities) < self.config.cull_alpha_thresh).squeeze()
        if extra_cull_mask is not None:
            culls = culls | extra_cull_mask
        if self.step > self.config.refine_every * self.config.reset_alpha_every:
            toobigs = (torch.exp(self.scales).max(dim=-1).values > self.config.cull_scale_thresh).squeeze()
            if self.step < self.config.stop_screen_size_at:
                assert self.max_2Dsize is not None
                toobigs = toobigs | (self.max_2Dsize > self.config.cull_screen_size).squeeze()
            culls = culls | toobigs
        for name, param in self.gauss_params.items():
            self.gauss_params[name] = torch.nn.Parameter(param[~culls])
        return culls

    def split_gaussians(self, split_mask, samps):
        from gsplat.cuda_legacy._torch_impl import quat_to_rotmat

        n_splits = int(split_mask.sum().item())
        centered_samples = torch.randn((samps * n_splits, 3), device=self.device)
        scaled_samples = torch.exp(self.scales[split_mask].repeat(samps, 1)) * centered_samples
        quats = self.quats[split_mask] / self.quats[split_mask].norm(dim=-1, keepdim=True)
        rots = quat_to_rotmat(quats.repeat(samps, 1))
        rotated_samples = torch.bmm(rots, scaled_samples[..., None]).squeeze()
        new_means = rotated_samples + self.means[split_mask].repeat(samps, 1)
        size_fac = 1.6
        new_scales = torch.log(torch.exp(self.scales[split_mask]) / size_fac).repeat(samps, 1)
        self.scales[split_mask] = torch.log(torch.exp(self.scales[split_mask]) / size_fac)
        out = {"means": new_means, "features_dc": self.features_dc[split_mask].repeat(samps, 1),
               "features_rest": self.features_rest[split_mask].repeat(samps, 1, 1),
               "opacities": self.opacities[split_mask].repeat(samps, 1), "scales": new_scales,
               "quats": self.quats[split_mask].repeat(samps, 1)}
        for name, param in self.gauss_params.items():
            if name not in out:
                out[name] = param[split_mask].repeat(samps, 1)
        return out

    def dup_gaussians(self, dup_mask):
        return {name: param[dup_mask] for name, param in self.gauss_params.items()}
