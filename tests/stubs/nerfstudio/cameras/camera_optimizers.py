from dataclasses import dataclass, field
from typing import Literal, Type

import torch


class CameraOptimizer(torch.nn.Module):
    """mode="off" only (the reference's setting, dn_model.py:128-130): cameras pass through unchanged."""

    def __init__(self, config, num_cameras: int, device, **kwargs):
        super().__init__()
        self.config, self.num_cameras = config, num_cameras
        if config.mode != "off":
            raise NotImplementedError("stub: camera optimizer modes other than 'off'")

    def apply_to_camera(self, camera):
        return camera.camera_to_worlds

    def get_loss_dict(self, loss_dict: dict) -> None:
        pass

    def get_metrics_dict(self, metrics_dict: dict) -> None:
        pass

    def get_param_groups(self, param_groups: dict) -> None:
        pass


@dataclass
class CameraOptimizerConfig:
    _target: Type = field(default_factory=lambda: CameraOptimizer)
    mode: Literal["off", "SO3xR3", "SE3"] = "off"

    def setup(self, **kwargs):
        return self._target(self, **kwargs)
