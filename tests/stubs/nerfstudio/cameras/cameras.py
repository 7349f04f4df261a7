from typing import Dict, Optional

import torch
from torch import Tensor


class Cameras:
    """The part of nerfstudio.cameras.cameras.Cameras that DNSplatterModel.get_outputs touches (dn_model.py:469-671)."""

    def __init__(self, camera_to_worlds: Tensor, fx: Tensor, fy: Tensor, cx: Tensor, cy: Tensor, width, height,
                 metadata: Optional[Dict] = None):
        self.camera_to_worlds = camera_to_worlds  # [B, 3, 4] (OpenGL convention, as nerfstudio stores it)
        B = camera_to_worlds.shape[0]
        dev = camera_to_worlds.device
        as_t = lambda v, dt: (v if isinstance(v, Tensor) else torch.full((B, 1), v)).to(dev, dt).reshape(B, 1)  # noqa: E731
        self.fx, self.fy = as_t(fx, torch.float32), as_t(fy, torch.float32)
        self.cx, self.cy = as_t(cx, torch.float32), as_t(cy, torch.float32)
        self.width, self.height = as_t(width, torch.int64), as_t(height, torch.int64)
        self.metadata = metadata

    @property
    def shape(self):
        return self.camera_to_worlds.shape[:-2]

    @property
    def device(self):
        return self.camera_to_worlds.device

    def rescale_output_resolution(self, scaling_factor) -> None:
        s = float(scaling_factor)
        self.fx, self.fy, self.cx, self.cy = self.fx * s, self.fy * s, self.cx * s, self.cy * s
        self.height = torch.floor(self.height * s + 0.5).to(torch.int64)
        self.width = torch.floor(self.width * s + 0.5).to(torch.int64)

    def get_intrinsics_matrices(self) -> Tensor:
        B = self.camera_to_worlds.shape[0]
        K = torch.zeros((B, 3, 3), dtype=torch.float32, device=self.device)
        K[:, 0, 0], K[:, 1, 1] = self.fx.squeeze(-1), self.fy.squeeze(-1)
        K[:, 0, 2], K[:, 1, 2] = self.cx.squeeze(-1), self.cy.squeeze(-1)
        K[:, 2, 2] = 1.0
        return K

    def to(self, device):
        c = Cameras(self.camera_to_worlds.to(device), self.fx, self.fy, self.cx, self.cy, self.width, self.height,
                    self.metadata)
        return c
