"""Forward-only render loops: several cameras per rasterization call.

The reference renders evaluation / dataset images one camera at a time through `model.get_outputs(camera)` under
`torch.no_grad()` (ns-eval / ns-render via /root/reference/eval_utils/rendering_evaluation.py:3-19,
dn_splatter/utils/utils.py:331-441 `gs_render_dataset_images`, scripts/render_video.py): per view one projection, two
binnings, two sorts, two compositing passes and ~25 torch glue launches.  `rasterization()` treats cameras as an
independent leading dimension, so `render_views` hands it `chunk` cameras at once: one
projection / binning / sort / compositing launch sequence for the whole chunk, both colour sets (RGB + expected depth and
the per-camera Gaussian normals) in one walk, and the image glue (background blend, clamp, depth fill, normal map) by the
fused kernels of compose.py per view.  Same output dict per view as `get_outputs` in eval mode."""
from __future__ import annotations

from typing import Dict, Iterable, Iterator, List, Sequence

import torch
from torch import Tensor

from .compose import compose_rgbd, normal_map
from .gaussians import gaussian_normals


@torch.no_grad()
def render_views(model, views: Sequence[int], chunk: int = 4) -> Iterator[Dict[str, Tensor]]:
    """`model`: a dn_step.DNSplatterStep (parameters + scene cameras).  Yields, in the order of `views`, dicts with
    `rgb [H,W,3]`, `depth [H,W,1]`, `normal [H,W,3]`, `accumulation [H,W,1]`, `background [3]`, `view`."""
    from .gsplat import rasterization_from_params

    sc, cfg = model.scene, model.config
    if model.device.type != "cuda":
        raise RuntimeError("render_views needs a CUDA model (no CPU fallback)")
    W, H = sc.width, sc.height
    sh = min(model.step // cfg.sh_degree_interval, cfg.sh_degree)
    opac = torch.sigmoid(model.opacities).squeeze(-1)
    views = list(views)
    for lo in range(0, len(views), max(1, int(chunk))):
        ids = views[lo:lo + max(1, int(chunk))]
        idx = torch.as_tensor(ids, device=model.device)
        viewmats, Ks, c2w = sc.viewmats[idx], sc.Ks[idx], sc.c2w[idx]
        # per-camera Gaussian normals (they face the camera and live in its frame, dn_model.py:617-636)
        normals = torch.stack([gaussian_normals(model.quats, model.scales, model.means, c2w[i])[0]
                               for i in range(len(ids))])
        render, alpha, info = rasterization_from_params(
            model.means, model.quats, model.scales, opac, model.features_dc, model.features_rest, viewmats=viewmats,
            Ks=Ks, width=W, height=H, sh_degree=sh, near_plane=0.01, far_plane=1e10, tile_size=16,
            render_mode="RGB+ED", absgrad=False, colors_b=normals)
        for i, v in enumerate(ids):
            rgb, depth = compose_rgbd(render[i:i + 1], alpha[i:i + 1], model.background)
            yield {"view": v, "rgb": rgb, "depth": depth, "normal": normal_map(info["render_b"][i]),
                   "accumulation": alpha[i], "background": model.background}


@torch.no_grad()
def render_dataset(model, views: Iterable[int], chunk: int = 4) -> List[Dict[str, Tensor]]:
    """`gs_render_dataset_images` without the file writing: all views rendered, results kept on the device."""
    return list(render_views(model, list(views), chunk=chunk))
