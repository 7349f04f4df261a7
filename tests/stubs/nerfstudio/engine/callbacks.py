from dataclasses import dataclass
from enum import Enum, auto
from typing import Callable, Dict, List, Optional, Tuple


class TrainingCallbackLocation(Enum):
    BEFORE_TRAIN_ITERATION = auto()
    AFTER_TRAIN_ITERATION = auto()
    AFTER_TRAIN = auto()


@dataclass
class TrainingCallbackAttributes:
    optimizers: Optional[object] = None
    grad_scaler: Optional[object] = None
    pipeline: Optional[object] = None
    trainer: Optional[object] = None


class TrainingCallback:
    """nerfstudio.engine.callbacks.TrainingCallback: run `func(*args, **kwargs, step=step)` at the given locations,
    either every `update_every_num_iters` steps or at the listed `iters`."""

    def __init__(self, where_to_run: List[TrainingCallbackLocation], func: Callable,
                 update_every_num_iters: Optional[int] = None, iters: Optional[Tuple[int, ...]] = None,
                 args: Optional[List] = None, kwargs: Optional[Dict] = None):
        self.where_to_run = where_to_run
        self.func = func
        self.update_every_num_iters = update_every_num_iters
        self.iters = iters
        self.args = args if args is not None else []
        self.kwargs = kwargs if kwargs is not None else {}

    def run_callback(self, step: int) -> None:
        if self.update_every_num_iters is not None:
            if step % self.update_every_num_iters == 0:
                self.func(*self.args, **self.kwargs, step=step)
        elif self.iters is not None:
            if step in self.iters:
                self.func(*self.args, **self.kwargs, step=step)

    def run_callback_at_location(self, step: int, location: TrainingCallbackLocation) -> None:
        if location in self.where_to_run:
            self.run_callback(step=step)
