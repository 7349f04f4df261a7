"""Fused multi-tensor Adam behind `torch.optim.Optimizer`.

Plug point: /root/reference/dn_splatter/dn_config.py:36-75 names one `AdamOptimizerConfig(lr, eps=1e-15)` per
Gaussian parameter group, and nerfstudio builds one optimizer object per group.  `FusedAdam` keeps exactly the
state layout those callers mutate by hand after every densify / prune
(`remove_from_optim`, `dup_in_optim`, dn_model.py:149-170,441-445,1120-1152):
`optimizer.state[param] = {"step", "exp_avg", "exp_avg_sq"}` and `param_groups[0]["params"]`.

`FusedAdam.step()` updates its own group with one kernel launch; `fused_step([...])` updates several
optimizers (all six live Gaussian groups) with a single launch of the same kernel.
"""
from __future__ import annotations

import ctypes
from typing import Iterable, List

import torch

from ._abi import check, lib
from .ops import _stream, kernel_timer


def _launch(entries, beta1, beta2, eps):
    """entries: list of (param, grad, exp_avg, exp_avg_sq, lr, step)."""
    maxt = lib.fsb_adam_max_tensors()
    for lo in range(0, len(entries), maxt):
        chunk = entries[lo:lo + maxt]
        n = len(chunk)
        VP = ctypes.c_void_p * n
        p = VP(*[e[0].data_ptr() for e in chunk])
        g = VP(*[e[1].data_ptr() for e in chunk])
        m = VP(*[e[2].data_ptr() for e in chunk])
        v = VP(*[e[3].data_ptr() for e in chunk])
        numel = (ctypes.c_int64 * n)(*[e[0].numel() for e in chunk])
        lr = (ctypes.c_double * n)(*[float(e[4]) for e in chunk])
        step = (ctypes.c_int64 * n)(*[int(e[5]) for e in chunk])
        check(lib.fsb_adam_multi(n, ctypes.addressof(p), ctypes.addressof(g), ctypes.addressof(m),
                                 ctypes.addressof(v), ctypes.addressof(numel), ctypes.addressof(lr),
                                 ctypes.addressof(step), beta1, beta2, eps, _stream()), "fsb_adam_multi")


class FusedAdam(torch.optim.Optimizer):
    """torch.optim.Adam(lr, betas, eps, weight_decay=0, amsgrad=False) semantics on libfsb200's kernel."""

    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.0):
        if weight_decay != 0.0:
            raise ValueError("FusedAdam implements the reference's setting only: weight_decay=0")
        super().__init__(params, dict(lr=lr, betas=betas, eps=eps, weight_decay=weight_decay))

    def _collect(self) -> List[tuple]:
        out = []
        for group in self.param_groups:
            for p in group["params"]:
                if p.grad is None:
                    continue
                if not p.is_cuda or p.dtype != torch.float32:
                    raise RuntimeError("FusedAdam needs fp32 CUDA parameters (no CPU fallback)")
                if p.grad.is_sparse:
                    raise RuntimeError("FusedAdam does not support sparse gradients")
                st = self.state[p]
                if len(st) == 0:
                    st["step"] = torch.tensor(0.0)
                    st["exp_avg"] = torch.zeros_like(p, memory_format=torch.preserve_format)
                    st["exp_avg_sq"] = torch.zeros_like(p, memory_format=torch.preserve_format)
                if not (p.is_contiguous() and st["exp_avg"].is_contiguous() and st["exp_avg_sq"].is_contiguous()):
                    raise RuntimeError("FusedAdam needs contiguous parameters and state")
                st["step"] = st["step"] + 1  # CPU scalar tensor like torch.optim.Adam's default: no device sync
                step = int(st["step"])
                out.append((p.data, p.grad.contiguous(), st["exp_avg"], st["exp_avg_sq"], group["lr"], step,
                            group["betas"], group["eps"]))
        return out

    @torch.no_grad()
    def step(self, closure=None):
        loss = closure() if closure is not None else None
        fused_step([self])
        return loss


@torch.no_grad()
def fused_step(optimizers: Iterable[FusedAdam]) -> None:
    """One kernel launch for all parameters of all given optimizers that share betas / eps."""
    buckets = {}
    for opt in optimizers:
        for e in opt._collect():
            buckets.setdefault((e[6], e[7]), []).append(e[:6])
    ev = kernel_timer.start("adam_multi")
    for (betas, eps), entries in buckets.items():
        _launch(entries, betas[0], betas[1], eps)
    kernel_timer.stop(ev)
