"""Device-resident training-view feeder with the `next_train` contract of DNSplatterDataManager.

/root/reference/dn_splatter/dn_datamanager.py:96-148 serves one view per iteration: it pops the next index from
`train_unseen_cameras` (sequential, refilled when empty, :100-102), `deepcopy`s the cached CPU batch (:103), moves
image / mask / depths / normal to the device and bilinearly resizes depth and normal maps whose shape differs from the
image's (:104-137), every single step.  Once the render step takes ~1 ms that host work and its 3-4 H2D copies are what
the GPU waits for.  `ResidentViewFeeder` keeps the same observable behaviour and does the work once:

  * at construction every cached batch is moved to the device, masks get their trailing channel, depth / normal maps are
    resized to the image size (same bilinear, antialias-free resize: torch's interpolate, what
    torchvision.transforms.functional.resize does for tensors);
  * `next_train(step)` pops indices in the reference's order and returns `(camera, batch)` where `batch` is a fresh dict
    of the RESIDENT tensors (the model only re-binds dict keys, dn_model.py:700-715; nothing mutates the tensors);
  * with `world_size > 1` every rank walks the same index sequence and takes view `rank` of each group of `world_size`
    consecutive indices (the camera sharding of SURVEY.md §8e: the step's camera batch is dealt over the ranks).

The camera objects are passed through untouched (`cameras[i:i+1].to(device)`, `metadata["cam_idx"]`, :138-147), so the
feeder works with nerfstudio's `Cameras` as well as with the test stub.  No kernels here: plumbing."""
from __future__ import annotations

from typing import Dict, List, Optional, Sequence, Tuple

import torch
import torch.nn.functional as F
from torch import Tensor


def _resize_like(t: Tensor, hw) -> Tensor:
    """[h,w,C] -> [H,W,C], bilinear without antialiasing (TF.resize(..., antialias=None) on a tensor)."""
    if tuple(t.shape[:2]) == tuple(hw):
        return t
    return F.interpolate(t.permute(2, 0, 1)[None].float(), size=tuple(hw), mode="bilinear",
                         align_corners=False)[0].permute(1, 2, 0).to(t.dtype).contiguous()


class ResidentViewFeeder:
    def __init__(self, cached_train: Sequence[Dict[str, Tensor]], cameras, device="cuda", world_size: int = 1,
                 rank: int = 0, load_depths: bool = True, load_normals: bool = True):
        if world_size < 1 or not 0 <= rank < world_size:
            raise ValueError((world_size, rank))
        self.device = torch.device(device)
        self.cameras = cameras
        self.world_size, self.rank = int(world_size), int(rank)
        self.load_depths, self.load_normals = load_depths, load_normals
        self.batches: List[Dict[str, Tensor]] = [self._prepare(b) for b in cached_train]
        self.train_unseen_cameras = list(range(len(self.batches)))
        self.image_idx = 0

    def _prepare(self, data: Dict[str, Tensor]) -> Dict[str, Tensor]:
        out = dict(data)
        out["image"] = data["image"].to(self.device)
        hw = out["image"].shape[:2]
        if "mask" in data:
            m = data["mask"].to(self.device)
            out["mask"] = m[..., None] if m.dim() == 2 else m
        if self.load_depths:
            for k in ("sensor_depth", "mono_depth"):
                if k in data:
                    out[k] = _resize_like(data[k].to(self.device), hw)
        if self.load_normals:
            if "normal" not in data:
                raise AssertionError("load_normals: every cached batch needs a 'normal' map (dn_datamanager.py:128)")
            out["normal"] = _resize_like(data["normal"].to(self.device), hw)
        return out

    def _pop(self) -> int:
        idx = self.train_unseen_cameras.pop(0)
        if len(self.train_unseen_cameras) == 0:
            self.train_unseen_cameras = list(range(len(self.batches)))
        return idx

    def next_index(self) -> int:
        """The view this rank trains on in the next iteration (advances the shared sequence by world_size)."""
        mine = None
        for r in range(self.world_size):
            idx = self._pop()
            if r == self.rank:
                mine = idx
        self.image_idx = mine
        return mine

    def next_train(self, step: int) -> Tuple[object, Dict[str, Tensor]]:
        idx = self.next_index()
        camera = self.cameras[idx:idx + 1]
        if hasattr(camera, "to"):
            camera = camera.to(self.device)
        if getattr(camera, "metadata", None) is None:
            camera.metadata = {}
        camera.metadata["cam_idx"] = idx
        return camera, dict(self.batches[idx])

    def targets(self) -> Dict[int, Dict[str, Tensor]]:
        """{view: batch} of the resident tensors, the form GraphedDNSplatterStep takes."""
        return {i: b for i, b in enumerate(self.batches)}
