"""The DN-Splatter training iteration, as FusionSense configures it, on top of the B200 kernels.

nerfstudio / torchmetrics are not installed in this image, so `DNSplatterModel` itself cannot be instantiated
here; this module restates, method by method and with the same names, what one `Trainer.train_iteration`
executes for the `dn-splatter` method:

  get_outputs      /root/reference/dn_splatter/dn_model.py:469-671   (binary opacities, rasterization RGB+ED,
                                                                      per-Gaussian normals, legacy normals pass)
  get_loss_dict    /root/reference/dn_splatter/dn_model.py:673-925 + splatfacto base loss (0.8 L1 + 0.2 (1-SSIM))
  optimizers       /root/reference/dn_splatter/dn_config.py:36-75    (Adam eps=1e-15, exp-decay on means)
  after_train      nerfstudio splatfacto: xys_grad_norm / vis_counts / max_2Dsize accumulation

Everything that touches Gaussians or pixels goes through the `gsplat` entry points of this package, exactly the
calls the reference makes; bench.py and smoke() drive this class.
"""
from __future__ import annotations

import os
from dataclasses import dataclass, field
from typing import Dict, Optional

import torch
import torch.nn.functional as F
from torch import Tensor

from .losses import DepthLoss, DepthLossType, FusedSSIM, TVLoss
from .synthetic import Scene


@dataclass
class DNSplatterStepConfig:
    """Effective FusionSense values (configs/config.py:3-39 over DNSplatterModelConfig, dn_model.py:55-142)."""
    sh_degree: int = 3
    sh_degree_interval: int = 1000
    ssim_lambda: float = 0.2
    use_depth_loss: bool = True
    depth_loss_type: DepthLossType = DepthLossType.EdgeAwareLogL1
    depth_tolerance: float = 0.1
    sensor_depth_lambda: float = 0.2
    use_depth_smooth_loss: bool = True
    smooth_loss_lambda: float = 0.1
    use_normal_loss: bool = True
    use_normal_tv_loss: bool = True
    normal_lambda: float = 0.4
    normal_supervision: str = "mono"  # "mono" (FusionSense: DSINE/omnidata maps) or "depth" (dn_model.py:773-795)
    two_d_gaussians: bool = True
    use_binary_opacities: bool = True
    binary_opacities_threshold: float = 0.9
    warmup_length: int = 500
    reset_alpha_every: int = 30
    refine_every: int = 100
    stop_split_at: int = 10000
    # nerfstudio 1.1.3 SplatfactoModelConfig defaults that refinement reads
    densify_grad_thresh: float = 0.0008
    densify_size_thresh: float = 0.01
    n_split_samples: int = 2
    cull_alpha_thresh: float = 0.1
    cull_scale_thresh: float = 0.5
    cull_screen_size: float = 0.15
    split_screen_size: float = 0.05
    stop_screen_size_at: int = 4000
    continue_cull_post_densification: bool = True
    background_color: str = "white"
    rasterize_mode: str = "classic"
    lrs: Dict[str, float] = field(default_factory=lambda: {
        "means": 1.6e-4, "features_dc": 0.0025, "features_rest": 0.0025 / 20, "opacities": 0.05, "scales": 0.005,
        "quats": 0.001})
    means_lr_final: float = 1.6e-6
    means_lr_max_steps: int = 30000
    fused_optimizer: bool = True
    fused_losses: bool = True  # dn_regularizer_loss instead of the torch loss classes
    fused_glue: bool = True  # gaussian_normals / densify_stats kernels instead of the inline torch ops
    # get_outputs through rasterization_from_params / compose_rgbd / normal_map and the flatness term through
    # flatness_loss (compose.py): the activations, the SH concatenation and the image-space glue of
    # dn_model.py:566-574, :602-613, :655-656, :817-819 run inside our kernels instead of ~55 torch launches
    # (FSB_FUSED_OUTPUTS=0 switches it off for A/B runs)
    fused_outputs: bool = field(default_factory=lambda: os.environ.get("FSB_FUSED_OUTPUTS", "1") != "0")
    # bin only the (Gaussian, tile) pairs that can pass the alpha test somewhere in the tile (csrc/isect_reach.cu;
    # 35-49 % fewer list entries, same images and gradients: tests/test_gpu_prune_lists.py).  Needs fused_outputs.
    prune_lists: bool = field(default_factory=lambda: os.environ.get("FSB_PRUNE_LISTS", "1") == "1")
    # the RGB + depth pass and the legacy normals pass composited by ONE walk of one (pruned, union) list
    # (rasterization_from_params(colors_b=...), csrc/raster.cu): one binning, one sort, one forward and one backward
    # compositing kernel per iteration instead of two of each.  Needs fused_outputs and fused_glue.
    fused_passes: bool = field(default_factory=lambda: os.environ.get("FSB_FUSED_PASSES", "1") == "1")
    # get_metrics_dict (dn_model.py:927-1000) as part of every iteration, the way nerfstudio's Trainer calls it: PSNR /
    # MSE / depth metrics by one fused launch + one SSIM launch, left on the device (metrics.py); LPIPS needs backbone
    # weights that this image does not have and is left to a user-supplied callable (metrics.RGBMetrics)
    step_metrics: bool = field(default_factory=lambda: os.environ.get("FSB_STEP_METRICS", "0") == "1")
    overlap_normals_pass: bool = True  # captured step only: the normals pass runs on a second stream beside the RGB+ED pass


def _torch_normal_from_depth_image(depths, fx, fy, cx, cy, img_size, c2w, device, smooth=False):
    """Literal torch form of normal_utils.py:23-46 + camera_utils.py:69-144, used only when the step is driven on
    the CPU by the reference arm (the CUDA path calls utils/normal_utils.py -> fsb_normal_from_depth)."""
    W, H = img_size
    u, v = torch.meshgrid(torch.arange(W), torch.arange(H), indexing="xy")
    coords = (torch.stack((u, v), dim=-1) + 0.5).view(-1, 2).float().to(device)
    d = depths.reshape(-1).float()
    pts = torch.stack(((coords[:, 0] - cx) * d / fx, (coords[:, 1] - cy) * d / fy, d), dim=-1)
    pts = (pts @ torch.linalg.inv(c2w[..., :3, :3]) + c2w[..., :3, 3]).view(H, W, 3)
    n = torch.cross(pts[1:H - 1, 2:W] - pts[1:H - 1, 0:W - 2], pts[0:H - 2, 1:W - 1] - pts[2:H, 1:W - 1], dim=-1)
    n = F.normalize(n, p=2, dim=-1)
    return F.pad(n.permute(2, 0, 1), (1, 1, 1, 1), mode="constant").permute(1, 2, 0)


class DNSplatterStep:
    def __init__(self, scene: Scene, config: Optional[DNSplatterStepConfig] = None, device="cuda", step: int = 3000,
                 gsplat_module=None, torch_losses=None):
        """`gsplat_module`: whatever `import gsplat` resolves to for dn_model.py (default: this package's sm_100a
        implementation).  bench.py's reference arm passes the CPU oracle here, the way the reference would run on
        an installed gsplat.
        `torch_losses`: a module with plain-torch `SSIM`, `DepthLoss`, `TVLoss` classes (the reference's
        dn_splatter/losses.py semantics).  Needed only with `fused_losses=False` or on the CPU: this package holds
        no torch loss arithmetic of its own (tests and the CPU arm pass oracle/dn_losses_ref.py)."""
        self.config = config or DNSplatterStepConfig()
        self.device = torch.device(device)
        if gsplat_module is None:
            from . import gsplat as gsplat_module
        self._rasterization = gsplat_module.rasterization
        self._rasterization_from_params = getattr(gsplat_module, "rasterization_from_params", None)
        self._rasterize_gaussians = gsplat_module.rasterize_gaussians
        self._quat_to_rotmat = gsplat_module.quat_to_rotmat
        sc = scene.to(self.device)
        self.scene = sc
        P = torch.nn.Parameter
        self.gauss_params = {
            "means": P(sc.means.clone()), "scales": P(sc.scales.clone()), "quats": P(sc.quats.clone()),
            "features_dc": P(sc.features_dc.clone()), "features_rest": P(sc.features_rest.clone()),
            "opacities": P(sc.opacities.clone()),
        }
        self.step = step
        self.training = True
        if self.config.fused_losses and self.device.type == "cuda":
            self.ssim = FusedSSIM(data_range=1.0, kernel_size=11)
            self.depth_loss = DepthLoss(self.config.depth_loss_type)
            self.smooth_loss = DepthLoss(DepthLossType.TV)
            self.tv_loss = TVLoss()
        else:
            if torch_losses is None:
                raise RuntimeError("DNSplatterStep(fused_losses=False) / a CPU device needs `torch_losses` (a module "
                                   "with the reference's torch loss classes): fusionsense_b200 has no torch fallback")
            self.ssim = torch_losses.SSIM(data_range=1.0, kernel_size=11).to(self.device)
            self.depth_loss = torch_losses.DepthLoss(self.config.depth_loss_type)
            self.smooth_loss = torch_losses.DepthLoss(DepthLossType.TV)
            self.tv_loss = torch_losses.TVLoss()
        self.background = torch.ones(3, device=self.device)  # background_color = "white" (dn_model.py:141)
        self.xys_grad_norm = None
        self.vis_counts = None
        self.max_2Dsize = None
        self.add_mask = None
        self.num_train_data = sc.viewmats.shape[0]
        self.last_size = (sc.height, sc.width)
        self._build_optimizers()

    # ---- parameters / optimisers ------------------------------------------------------------
    def __getattr__(self, name):
        gp = self.__dict__.get("gauss_params")
        if gp is not None and name in gp:
            return gp[name]
        raise AttributeError(name)

    @property
    def num_points(self) -> int:
        return self.gauss_params["means"].shape[0]

    def _build_optimizers(self):
        from .optim import FusedAdam

        cfg = self.config
        self.optimizers = {}
        fused = cfg.fused_optimizer and self.device.type == "cuda"
        for name, lr in cfg.lrs.items():
            cls = FusedAdam if fused else torch.optim.Adam
            self.optimizers[name] = cls([self.gauss_params[name]], lr=lr, eps=1e-15)
        self._fused_optim = fused

    def _means_lr(self) -> float:
        cfg = self.config
        t = min(self.step / cfg.means_lr_max_steps, 1.0)
        return cfg.lrs["means"] * (cfg.means_lr_final / cfg.lrs["means"]) ** t

    # ---- dn_model.py:469-671 ----------------------------------------------------------------
    def get_outputs(self, cam_idx: int) -> Dict[str, Tensor]:
        cfg, sc = self.config, self.scene
        if cfg.use_binary_opacities and self.step > cfg.warmup_length:
            skip_steps = cfg.reset_alpha_every * cfg.refine_every
            if not self.step % skip_steps == 0 and self.step % skip_steps not in range(1, 200 + 1):
                # dn_model.py:497-503 assigns where(op >= thr, 1, 0) to .data; in place here (same values), so the
                # parameter keeps its storage and a captured graph keeps reading the live tensor
                self.opacities.data.ge_(cfg.binary_opacities_threshold)
        opacities_crop, means_crop = self.opacities, self.means
        scales_crop, quats_crop = self.scales, self.quats
        BLOCK_WIDTH = 16
        if isinstance(cam_idx, Tensor):  # device index (captured step: the view is chosen at replay time)
            # one gather for the three camera matrices: rows of [viewmat | K | c2w] (41 floats per view)
            if getattr(self, "_cam_pack", None) is None:
                V = sc.viewmats.shape[0]
                self._cam_pack = torch.cat([sc.viewmats.reshape(V, 16), sc.Ks.reshape(V, 9), sc.c2w.reshape(V, 16)],
                                           dim=1).contiguous()
            row = self._cam_pack.index_select(0, cam_idx)[0]
            viewmat, K, c2w = row[0:16].view(1, 4, 4), row[16:25].view(1, 3, 3), row[25:41].view(1, 4, 4)
        else:
            viewmat = sc.viewmats[cam_idx:cam_idx + 1]
            K = sc.Ks[cam_idx:cam_idx + 1]
            c2w = sc.c2w[cam_idx:cam_idx + 1]
        W, H = sc.width, sc.height
        self.last_size = (H, W)
        self._last_cam = cam_idx
        sh_degree_to_use = min(self.step // cfg.sh_degree_interval, cfg.sh_degree)
        if cfg.fused_outputs and self.device.type == "cuda" and self._rasterization_from_params is not None:
            return self._get_outputs_fused(viewmat, K, c2w, W, H, BLOCK_WIDTH, sh_degree_to_use)
        colors_crop = torch.cat((self.features_dc[:, None, :], self.features_rest), dim=1)
        render, alpha, info = self._rasterization(
            means=means_crop,
            quats=quats_crop / quats_crop.norm(dim=-1, keepdim=True),
            scales=torch.exp(scales_crop),
            opacities=torch.sigmoid(opacities_crop).squeeze(-1),
            colors=colors_crop,
            viewmats=viewmat,
            Ks=K,
            width=W,
            height=H,
            tile_size=BLOCK_WIDTH,
            packed=False,
            near_plane=0.01,
            far_plane=1e10,
            render_mode="RGB+ED",
            sh_degree=sh_degree_to_use,
            sparse_grad=False,
            absgrad=True,
            rasterize_mode=cfg.rasterize_mode,
        )
        if self.training and info["means2d"].requires_grad:
            info["means2d"].retain_grad()
        self.xys = info["means2d"]
        self.radii = info["radii"][0]
        self.depths = info["depths"]
        self.conics = info["conics"]
        self.num_tiles_hit = info["tiles_per_gauss"]
        background = self.background
        rgb = render[:, ..., :3] + (1 - alpha) * background
        rgb = torch.clamp(rgb, 0.0, 1.0)
        depth_im = render[:, ..., 3:4]
        depth_im = torch.where(alpha > 0, depth_im, depth_im.detach().max()).squeeze(0)

        # The normals pass needs only the projection outputs.  In a captured step (static-capacity mode: it bins on
        # its own, no host read) it is forked onto a second stream right after the projection, so its binning, sort
        # and compositing fill the SMs that the RGB+ED pass leaves idle in its kernel tails; autograd runs each
        # backward on its forward's stream, so the two raster backwards overlap the same way.
        proj_done = info.get("projection_done") if cfg.overlap_normals_pass else None
        main_stream = side_stream = None
        if proj_done is not None:
            main_stream = torch.cuda.current_stream()
            if getattr(self, "_normals_stream", None) is None:
                self._normals_stream = torch.cuda.Stream()
            side_stream = self._normals_stream
            side_stream.wait_event(proj_done)
            for t in (self.xys, self.depths, self.radii, self.conics, self.num_tiles_hit, c2w):
                t.record_stream(side_stream)
            torch.cuda.set_stream(side_stream)
        try:
            normals_im = self._normals_pass(quats_crop, scales_crop, means_crop, opacities_crop, c2w, H, W, BLOCK_WIDTH)
        finally:
            if side_stream is not None:
                torch.cuda.set_stream(main_stream)
        if side_stream is not None:
            main_stream.wait_stream(side_stream)
            normals_im.record_stream(main_stream)
        normals_im = normals_im / normals_im.norm(dim=-1, keepdim=True)
        normals_im = (normals_im + 1) / 2
        return {"rgb": rgb.squeeze(0), "depth": depth_im, "normal": normals_im, "accumulation": alpha.squeeze(0),
                "background": background}

    def _get_outputs_fused(self, viewmat, K, c2w, W, H, BLOCK_WIDTH, sh_degree_to_use) -> Dict[str, Tensor]:
        """get_outputs with the torch glue folded into the kernels: same values as the path above (the literal
        restatement of dn_model.py:566-671) to fp32 rounding, ~55 fewer launches per iteration.
        tests/test_gpu_step.py compares the two paths output by output and gradient by gradient."""
        from .compose import compose_rgbd, normal_map

        cfg = self.config
        opac = torch.sigmoid(self.opacities)  # [N,1]; the reference evaluates it twice (dn_model.py:574, :650)
        if cfg.fused_passes and cfg.fused_glue:
            from .gaussians import gaussian_normals

            normals, self.normals_world = gaussian_normals(self.quats, self.scales, self.means, c2w.squeeze(0))
            render, alpha, info = self._rasterization_from_params(
                self.means, self.quats, self.scales, opac.squeeze(-1), self.features_dc, self.features_rest,
                viewmats=viewmat, Ks=K, width=W, height=H, sh_degree=sh_degree_to_use, near_plane=0.01, far_plane=1e10,
                tile_size=BLOCK_WIDTH, render_mode="RGB+ED", absgrad=True, colors_b=normals)
            if self.training and info["means2d"].requires_grad:
                info["means2d"].retain_grad()
            self.xys = info["means2d"]
            self.radii = info["radii"][0]
            self.depths = info["depths"]
            self.conics = info["conics"]
            self.num_tiles_hit = info["tiles_per_gauss"]
            rgb, depth_im = compose_rgbd(render, alpha, self.background)
            normals_im = normal_map(info["render_b"].squeeze(0))  # squeeze: a view in backward too (select is not)
            return {"rgb": rgb, "depth": depth_im, "normal": normals_im, "accumulation": alpha.squeeze(0),
                    "background": self.background}
        render, alpha, info = self._rasterization_from_params(
            self.means, self.quats, self.scales, opac.squeeze(-1), self.features_dc, self.features_rest,
            viewmats=viewmat, Ks=K, width=W, height=H, sh_degree=sh_degree_to_use, near_plane=0.01, far_plane=1e10,
            tile_size=BLOCK_WIDTH, render_mode="RGB+ED", absgrad=True, **({"prune_lists": True} if cfg.prune_lists else {}))
        if self.training and info["means2d"].requires_grad:
            info["means2d"].retain_grad()
        self.xys = info["means2d"]
        self.radii = info["radii"][0]
        self.depths = info["depths"]
        self.conics = info["conics"]
        self.num_tiles_hit = info["tiles_per_gauss"]
        background = self.background
        proj_done = info.get("projection_done") if cfg.overlap_normals_pass else None
        main_stream = side_stream = None
        if proj_done is not None:
            main_stream = torch.cuda.current_stream()
            if getattr(self, "_normals_stream", None) is None:
                self._normals_stream = torch.cuda.Stream()
            side_stream = self._normals_stream
            side_stream.wait_event(proj_done)
            for t in (self.xys, self.depths, self.radii, self.conics, self.num_tiles_hit, c2w, opac):
                t.record_stream(side_stream)
            torch.cuda.set_stream(side_stream)
        try:
            normals_im = self._normals_pass(self.quats, self.scales, self.means, self.opacities, c2w, H, W,
                                            BLOCK_WIDTH, opac=opac)
        finally:
            if side_stream is not None:
                torch.cuda.set_stream(main_stream)
        # the RGB / depth composition runs on the main stream beside the normals pass
        rgb, depth_im = compose_rgbd(render, alpha, background)
        if side_stream is not None:
            main_stream.wait_stream(side_stream)
            normals_im.record_stream(main_stream)
        normals_im = normal_map(normals_im)
        return {"rgb": rgb, "depth": depth_im, "normal": normals_im, "accumulation": alpha.squeeze(0),
                "background": background}

    def _normals_pass(self, quats_crop, scales_crop, means_crop, opacities_crop, c2w, H, W, BLOCK_WIDTH, opac=None):
        """dn_model.py:617-653: per-Gaussian normals, then the legacy `rasterize_gaussians` pass over them."""
        cfg = self.config
        if cfg.fused_glue and self.device.type == "cuda":
            from .gaussians import gaussian_normals

            normals, self.normals_world = gaussian_normals(quats_crop, scales_crop, means_crop, c2w.squeeze(0))
        else:
            quats_n = quats_crop / quats_crop.norm(dim=-1, keepdim=True)
            normals = F.one_hot(torch.argmin(scales_crop, dim=-1), num_classes=3).float()
            rots = self._quat_to_rotmat(quats_n)
            normals = torch.bmm(rots, normals[:, :, None]).squeeze(-1)
            normals = F.normalize(normals, dim=1)
            viewdirs = -means_crop.detach() + c2w.detach()[..., :3, 3]
            viewdirs = viewdirs / viewdirs.norm(dim=-1, keepdim=True)
            dots = (normals * viewdirs).sum(-1)
            negative_dot_indices = dots < 0
            normals = torch.where(negative_dot_indices[:, None], -normals, normals)
            self.normals_world = normals.detach()
            normals = normals @ c2w.squeeze(0)[:3, :3]
        xys = self.xys[0, ...].detach()
        return self._rasterize_gaussians(xys, self.depths[0, ...], self.radii, self.conics[0, ...],
                                         self.num_tiles_hit[0, ...], normals,
                                         torch.sigmoid(opacities_crop) if opac is None else opac, H, W, BLOCK_WIDTH)

    # ---- splatfacto base loss + dn_model.py:673-925 -----------------------------------------
    def get_loss_dict(self, outputs, batch) -> Dict[str, Tensor]:
        cfg = self.config
        gt_rgb = batch["image"]  # already composited / on device
        pred_img = outputs["rgb"]
        fused = cfg.fused_losses and self.device.type == "cuda"
        combine = fused and cfg.fused_outputs  # main_loss assembled by one launch (compose.combine_losses)
        ssim_val = self.ssim(gt_rgb.permute(2, 0, 1)[None, ...], pred_img.permute(2, 0, 1)[None, ...])
        simloss = None if combine else 1 - ssim_val
        if combine:
            rgb_loss = None
        elif fused:
            rgb_loss = cfg.ssim_lambda * simloss  # the L1 half rides in the fused kernel below
        else:
            Ll1 = torch.abs(gt_rgb - pred_img).mean()
            rgb_loss = (1 - cfg.ssim_lambda) * Ll1 + cfg.ssim_lambda * simloss

        depth_out = outputs["depth"]
        sensor_depth_gt = batch["sensor_depth"]
        gt_normal = self._gt_normal(batch, depth_out) if cfg.use_normal_loss else None
        if fused:
            from .losses import dn_regularizer_loss

            reg = dn_regularizer_loss(
                depth_out, sensor_depth_gt, batch["image"], outputs["normal"], gt_normal,
                pred_rgb=pred_img, gt_rgb=gt_rgb, rgb_l1_lambda=1 - cfg.ssim_lambda,
                depth_tolerance=cfg.depth_tolerance,
                sensor_depth_lambda=cfg.sensor_depth_lambda if cfg.use_depth_loss else 0.0,
                smooth_loss_lambda=cfg.smooth_loss_lambda if cfg.use_depth_smooth_loss else 0.0,
                normal_l1_lambda=cfg.normal_lambda if cfg.use_normal_loss else 0.0,
                normal_tv_lambda=cfg.normal_lambda if (cfg.use_normal_loss and cfg.use_normal_tv_loss) else 0.0)
            depth_loss, normal_loss = reg, 0
        else:
            gt_img = batch["image"].clamp(min=10 / 255.0)
            depth_loss = 0
            if cfg.use_depth_loss and cfg.sensor_depth_lambda > 0.0:
                valid_gt_mask = sensor_depth_gt > cfg.depth_tolerance
                depth_loss = depth_loss + cfg.sensor_depth_lambda * self.depth_loss(
                    depth_out, sensor_depth_gt.float(), gt_img, valid_gt_mask)
            if cfg.use_depth_smooth_loss:
                depth_loss = depth_loss + cfg.smooth_loss_lambda * self.smooth_loss(depth_out)
            normal_loss = 0
            if cfg.use_normal_loss:
                pred_normal = outputs["normal"]
                normal_loss = normal_loss + torch.abs(gt_normal - pred_normal).mean()
                if cfg.use_normal_tv_loss:
                    normal_loss = normal_loss + self.tv_loss(pred_normal)
        if combine:
            from .compose import combine_losses, flatness_loss

            flat = flatness_loss(self.scales) if cfg.two_d_gaussians else None
            main_loss = combine_losses(ssim_val, reg, flat, cfg.ssim_lambda, cfg.normal_lambda)
            return {"main_loss": main_loss, "scale_reg": torch.zeros((), device=self.device)}
        if cfg.two_d_gaussians:
            normal_loss = normal_loss + torch.min(torch.exp(self.scales), dim=1, keepdim=True)[0].mean()
        main_loss = rgb_loss + depth_loss + cfg.normal_lambda * normal_loss
        return {"main_loss": main_loss, "scale_reg": torch.zeros((), device=self.device)}

    def _gt_normal(self, batch, depth_out) -> Tensor:
        """dn_model.py:770-795: monocular normal maps from the batch, or pseudo normals from the rendered depth."""
        cfg, sc = self.config, self.scene
        if "normal" in batch and cfg.normal_supervision == "mono":
            return batch["normal"]
        if cfg.normal_supervision == "depth":
            if self.device.type == "cuda":
                from .utils.normal_utils import normal_from_depth_image
            else:
                normal_from_depth_image = _torch_normal_from_depth_image  # reference arm (oracle-driven, CPU)
            if isinstance(self._last_cam, Tensor):
                raise NotImplementedError("normal_supervision='depth' reads the intrinsics on the host "
                                          "(dn_model.py:781-784 .item()); not available in a captured step")
            K = sc.Ks[self._last_cam]
            gt_normal = normal_from_depth_image(
                depths=depth_out.detach(), fx=K[0, 0].item(), fy=K[1, 1].item(), cx=K[0, 2].item(), cy=K[1, 2].item(),
                img_size=(sc.width, sc.height), c2w=torch.eye(4, dtype=torch.float, device=depth_out.device),
                device=self.device, smooth=False)
            gt_normal = gt_normal @ torch.diag(torch.tensor([1, -1, -1], device=depth_out.device, dtype=depth_out.dtype))
            return (1 + gt_normal) / 2
        raise RuntimeError("normal supervision with monocular normals enabled but the batch holds none "
                           "(dn_model.py:796-803 quits here)")

    # ---- dn_model.py:927-1000 -------------------------------------------------------------------
    @torch.no_grad()
    def get_metrics_dict(self, outputs, batch) -> Dict[str, Tensor]:
        """rgb_mse / rgb_psnr / rgb_ssim / depth_* as device scalars (no float(): nothing synchronises) plus
        gaussian_count; `rgb_lpips` is absent (no backbone weights offline)."""
        from .metrics import step_metrics

        if getattr(self, "_metrics_ssim", None) is None:
            self._metrics_ssim = FusedSSIM(data_range=1.0, kernel_size=11)
            self._metrics_out = torch.empty((10,), dtype=torch.float32, device=self.device)
        d = step_metrics(outputs, batch, self.config.depth_tolerance, ssim=self._metrics_ssim, out=self._metrics_out)
        d["gaussian_count"] = self.num_points
        return d

    # ---- splatfacto after_train ---------------------------------------------
    @torch.no_grad()
    def after_train(self, skip_flag=None):
        if self.step >= self.config.stop_split_at:
            return
        if self.config.fused_glue and self.device.type == "cuda":
            from .gaussians import densify_stats

            if self.xys_grad_norm is None:
                self.xys_grad_norm = torch.zeros(self.num_points, device=self.device, dtype=torch.float32)
                self.vis_counts = torch.ones(self.num_points, device=self.device, dtype=torch.float32)
            if self.max_2Dsize is None:
                self.max_2Dsize = torch.zeros(self.num_points, device=self.device, dtype=torch.float32)
            densify_stats(self.radii, self.xys.absgrad[0], float(max(self.last_size[0], self.last_size[1])),
                          self.xys_grad_norm, self.vis_counts, self.max_2Dsize, skip_flag=skip_flag)
            return
        visible_mask = (self.radii > 0).flatten()
        grads = self.xys.absgrad[0][visible_mask].norm(dim=-1)
        if self.xys_grad_norm is None:
            self.xys_grad_norm = torch.zeros(self.num_points, device=self.device, dtype=torch.float32)
            self.vis_counts = torch.ones(self.num_points, device=self.device, dtype=torch.float32)
        self.vis_counts[visible_mask] += 1
        self.xys_grad_norm[visible_mask] += grads
        if self.max_2Dsize is None:
            self.max_2Dsize = torch.zeros_like(self.radii, dtype=torch.float32)
        newradii = self.radii.detach()[visible_mask]
        self.max_2Dsize[visible_mask] = torch.maximum(
            self.max_2Dsize[visible_mask], newradii / float(max(self.last_size[0], self.last_size[1])))

    # ---- dn_model.py:326-451 / :1249-1276 (AFTER_TRAIN_ITERATION callbacks, every refine_every steps) ----------
    def refinement_after(self, samples=None):
        from .densify import refinement_after

        return refinement_after(self, self.optimizers, self.step, samples=samples)

    def hull_pruning(self, visual_hull: Tensor, scale_factor: float = 1.0):
        from .densify import hull_pruning

        return hull_pruning(self, self.optimizers, self.step, visual_hull, scale_factor)

    # ---- Trainer.train_iteration ------------------------------------------------------------
    def train_iteration(self, cam_idx: int, batch: Dict[str, Tensor]) -> Tensor:
        for opt in self.optimizers.values():
            opt.zero_grad(set_to_none=True)
        outputs = self.get_outputs(cam_idx)
        if self.config.step_metrics:
            self.last_metrics = self.get_metrics_dict(outputs, batch)
        loss_dict = self.get_loss_dict(outputs, batch)
        loss = loss_dict["main_loss"] + loss_dict["scale_reg"]
        loss.backward()
        self.optimizers["means"].param_groups[0]["lr"] = self._means_lr()
        self.optimizer_step()
        self.after_train()
        self.step += 1
        return loss.detach()

    def optimizer_step(self):
        """optimizers.optimizer_scaler_step_some: every Gaussian group steps every iteration."""
        if self._fused_optim:
            from .optim import fused_step

            fused_step(self.optimizers.values())  # one launch for all six groups
        else:
            for opt in self.optimizers.values():
                opt.step()

    @torch.no_grad()
    def render_targets(self, cam_idx: int, perturb: float = 0.02, seed: int = 0) -> Dict[str, Tensor]:
        """Ground truth for a synthetic view: a render of a perturbed copy of the scene."""
        g = torch.Generator(device="cpu").manual_seed(seed + cam_idx)
        saved = {k: v.data.clone() for k, v in self.gauss_params.items()}
        for k, v in self.gauss_params.items():
            noise = torch.randn(v.shape, generator=g).to(self.device)
            v.data.add_(perturb * v.data.abs().mean() * noise)
        ub, self.config.use_binary_opacities, self.training = self.config.use_binary_opacities, False, False
        out = self.get_outputs(cam_idx)
        self.config.use_binary_opacities, self.training = ub, True
        for k, v in self.gauss_params.items():
            v.data.copy_(saved[k])
        depth = torch.where(out["accumulation"] > 0.5, out["depth"], torch.zeros_like(out["depth"]))
        return {"image": out["rgb"].contiguous(), "sensor_depth": depth.contiguous(), "normal": out["normal"].contiguous()}
