from dataclasses import dataclass, field
from typing import Dict, List, Optional, Tuple, Type

import torch


@dataclass
class OptimizerConfig:
    _target: Type = torch.optim.Adam
    lr: float = 0.0005
    eps: float = 1e-08
    max_norm: Optional[float] = None

    def setup(self, params):
        kwargs = {k: v for k, v in vars(self).items() if k not in ("_target", "max_norm")}
        return self._target(params, **kwargs)


@dataclass
class AdamOptimizerConfig(OptimizerConfig):
    _target: Type = torch.optim.Adam
    weight_decay: float = 0


class Optimizers:
    """nerfstudio.engine.optimizers.Optimizers: one optimizer object per parameter group (name -> [params])."""

    def __init__(self, config: Dict[str, Dict], param_groups: Dict[str, List[torch.nn.Parameter]]):
        self.config = config
        self.optimizers: Dict[str, torch.optim.Optimizer] = {}
        self.schedulers = {}
        self.parameters = {}
        for name, params in param_groups.items():
            self.optimizers[name] = config[name]["optimizer"].setup(params=params)
            self.parameters[name] = params

    def zero_grad_all(self):
        for opt in self.optimizers.values():
            opt.zero_grad(set_to_none=True)

    def optimizer_step_all(self):
        for opt in self.optimizers.values():
            opt.step()
