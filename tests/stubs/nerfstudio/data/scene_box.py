from dataclasses import dataclass

import torch


@dataclass
class OrientedBox:
    R: torch.Tensor
    T: torch.Tensor
    S: torch.Tensor

    def within(self, pts):
        local = (pts - self.T) @ self.R
        return ((local.abs() <= self.S / 2).all(dim=-1))[..., None]


@dataclass
class SceneBox:
    aabb: torch.Tensor
