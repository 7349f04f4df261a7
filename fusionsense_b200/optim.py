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


class CapturedAdam:
    """The same update for a CUDA-graph-captured step (fsb_adam_multi_dev).

    A captured launch freezes its by-value arguments, so the per-step scalars (lr / (1 - beta1^t), sqrt(1 - beta2^t))
    live in a small device buffer that `advance()` refreshes before every replay with fsb_upload_small (the bytes
    ride in kernel arguments: no pinned staging, no synchronisation).  `advance()` also moves the optimizers'
    `state[p]["step"]` forward, so the state stays what torch.optim.Adam would hold; `rollback(k)` undoes k steps
    that the device skipped (skip_flag raised by an overflowed static-capacity step)."""

    def __init__(self, optimizers: Iterable[FusedAdam]):
        self.entries = []  # (optimizer, group, param)
        for opt in optimizers:
            for group in opt.param_groups:
                for p in group["params"]:
                    self.entries.append((opt, group, p))
        maxt = lib.fsb_adam_max_tensors()
        if not 0 < len(self.entries) <= maxt:
            raise RuntimeError(f"CapturedAdam handles 1..{maxt} parameter tensors, got {len(self.entries)}")
        g0 = self.entries[0][1]
        self.betas, self.eps = g0["betas"], g0["eps"]
        for opt, group, p in self.entries:
            if group["betas"] != self.betas or group["eps"] != self.eps:
                raise RuntimeError("CapturedAdam needs one (betas, eps) setting for all groups")
            if not p.is_cuda or p.dtype != torch.float32 or not p.is_contiguous():
                raise RuntimeError("CapturedAdam needs contiguous fp32 CUDA parameters (no CPU fallback)")
            st = opt.state[p]
            if len(st) == 0:
                st["step"] = torch.tensor(0.0)
                st["exp_avg"] = torch.zeros_like(p, memory_format=torch.preserve_format)
                st["exp_avg_sq"] = torch.zeros_like(p, memory_format=torch.preserve_format)
        self.maxt = maxt
        self.hyper = torch.zeros(2 * maxt, dtype=torch.float32, device=self.entries[0][2].device)
        self._host = (ctypes.c_float * (2 * maxt))()

    @torch.no_grad()
    def launch_xchg(self, exchange, skip_flag=None) -> None:
        """Multi-GPU: the same update with the gradient gathered from the ranks' reduced slices in peer memory
        (fsb_adam_multi_xchg, csrc/grad_exchange.cu); `exchange` = a dist.PeerGradExchange whose exchange() has run
        on this step's gradients, in the order of `self.entries`."""
        n = len(self.entries)
        off, world, r_ptrs, S = exchange.adam_args()
        VP = ctypes.c_void_p * n
        ps, ms, vs, ns = [], [], [], []
        for opt, group, p in self.entries:
            st = opt.state[p]
            ps.append(p.data_ptr()); ms.append(st["exp_avg"].data_ptr()); vs.append(st["exp_avg_sq"].data_ptr())
            ns.append(p.numel())
        if ns != list(exchange.ns):
            raise RuntimeError("CapturedAdam.launch_xchg: the exchange was run on a different tensor list")
        p_, m_, v_ = VP(*ps), VP(*ms), VP(*vs)
        numel = (ctypes.c_int64 * n)(*ns)
        offs = (ctypes.c_int64 * n)(*off[:n])
        check(lib.fsb_adam_multi_xchg(n, ctypes.addressof(p_), ctypes.addressof(m_), ctypes.addressof(v_),
                                      ctypes.addressof(numel), ctypes.addressof(offs), world,
                                      ctypes.addressof(r_ptrs), S, self.hyper.data_ptr(), self.maxt,
                                      None if skip_flag is None else skip_flag.data_ptr(), self.betas[0],
                                      self.betas[1], self.eps, _stream()), "fsb_adam_multi_xchg")

    @torch.no_grad()
    def launch_range(self, exchange, flat_begin: int, flat_end: int, skip_flag=None) -> None:
        """The update of the elements whose index in the exchange's flat gradient buffer lies in [flat_begin, flat_end)
        (multiples of 4), the gradient read from this rank's own, all-reduced buffer: one piece of
        PeerGradExchange.exchange_and_adam's pipeline.  Tensors outside the range ride along with n = 0, so every
        tensor keeps its slot in the hyper-parameter table."""
        n = len(self.entries)
        off = exchange.off
        VP = ctypes.c_void_p * n
        ps, ms, vs, ns, offs = [], [], [], [], []
        for i, (opt, group, p) in enumerate(self.entries):
            st = opt.state[p]
            if p.numel() != exchange.ns[i]:
                raise RuntimeError("CapturedAdam.launch_range: the exchange was run on a different tensor list")
            lo = min(max(0, flat_begin - off[i]), p.numel())
            hi = min(p.numel(), max(0, flat_end - off[i]))
            cnt = max(0, hi - lo)
            ps.append(p.data_ptr() + 4 * lo); ms.append(st["exp_avg"].data_ptr() + 4 * lo)
            vs.append(st["exp_avg_sq"].data_ptr() + 4 * lo)
            ns.append(cnt); offs.append(off[i] + lo)
        p_, m_, v_ = VP(*ps), VP(*ms), VP(*vs)
        numel = (ctypes.c_int64 * n)(*ns)
        offs_ = (ctypes.c_int64 * n)(*offs)
        g_self = (ctypes.c_void_p * 1)(int(exchange.G.data_ptr()))
        check(lib.fsb_adam_multi_xchg(n, ctypes.addressof(p_), ctypes.addressof(m_), ctypes.addressof(v_),
                                      ctypes.addressof(numel), ctypes.addressof(offs_), 1, ctypes.addressof(g_self),
                                      exchange.total, self.hyper.data_ptr(), self.maxt,
                                      None if skip_flag is None else skip_flag.data_ptr(), self.betas[0],
                                      self.betas[1], self.eps, _stream()), "fsb_adam_multi_xchg")

    @torch.no_grad()
    def launch(self, skip_flag=None) -> None:
        """Inside the capture, after backward(): one launch for all tensors, scalars read from `self.hyper`."""
        n = len(self.entries)
        VP = ctypes.c_void_p * n
        ps, gs, ms, vs, ns = [], [], [], [], []
        for opt, group, p in self.entries:
            if p.grad is None:
                raise RuntimeError("CapturedAdam.launch: a parameter has no gradient in the captured step")
            st = opt.state[p]
            ps.append(p.data_ptr()); gs.append(p.grad.contiguous().data_ptr())
            ms.append(st["exp_avg"].data_ptr()); vs.append(st["exp_avg_sq"].data_ptr()); ns.append(p.numel())
        p_, g_, m_, v_ = VP(*ps), VP(*gs), VP(*ms), VP(*vs)
        numel = (ctypes.c_int64 * n)(*ns)
        check(lib.fsb_adam_multi_dev(n, ctypes.addressof(p_), ctypes.addressof(g_), ctypes.addressof(m_),
                                     ctypes.addressof(v_), ctypes.addressof(numel), self.hyper.data_ptr(),
                                     None if skip_flag is None else skip_flag.data_ptr(), self.betas[0],
                                     self.betas[1], self.eps, _stream()), "fsb_adam_multi_dev")

    def advance(self) -> None:
        """Before every replay: step counts += 1, then upload this step's scalars (stream-ordered)."""
        b1, b2 = self.betas
        for i, (opt, group, p) in enumerate(self.entries):
            st = opt.state[p]
            st["step"] = st["step"] + 1
            t = float(st["step"])
            self._host[i] = group["lr"] / (1.0 - b1 ** t)
            self._host[self.maxt + i] = (1.0 - b2 ** t) ** 0.5
        check(lib.fsb_upload_small(self.hyper.data_ptr(), ctypes.addressof(self._host), 8 * self.maxt, _stream()),
              "fsb_upload_small")

    def rollback(self, k: int) -> None:
        for opt, group, p in self.entries:
            opt.state[p]["step"] = opt.state[p]["step"] - k


# ---------------------------------------------------------------------------------------------
# nerfstudio plug point
# ---------------------------------------------------------------------------------------------
from dataclasses import dataclass as _dataclass  # noqa: E402
from typing import Optional as _Optional  # noqa: E402
from typing import Type as _Type  # noqa: E402


@_dataclass
class FusedAdamOptimizerConfig:
    """Drop-in for `nerfstudio.engine.optimizers.AdamOptimizerConfig` in the method config
    (/root/reference/dn_splatter/dn_config.py:36-75 names one `AdamOptimizerConfig(lr=..., eps=1e-15)` per Gaussian
    parameter group): same fields, same `setup(params)` protocol, `_target` = FusedAdam.  nerfstudio's `Optimizers`
    then builds one FusedAdam per group; `FusedAdam.step()` is one launch per group, `fused_step(optimizers)` /
    `CapturedAdam` update all groups with one.  `use_fused_adam(method_config)` swaps it into an existing config."""

    _target: _Type = FusedAdam
    lr: float = 0.0005
    eps: float = 1e-08
    max_norm: _Optional[float] = None
    weight_decay: float = 0

    def setup(self, params) -> FusedAdam:
        if self.max_norm is not None:
            raise ValueError("FusedAdamOptimizerConfig: gradient clipping (max_norm) is not implemented "
                             "(the reference does not use it)")
        return self._target(params, lr=self.lr, eps=self.eps, weight_decay=self.weight_decay)


def use_fused_adam(optimizers_config: dict, groups=("means", "features_dc", "features_rest", "opacities", "scales",
                                                     "quats")) -> dict:
    """`optimizers` dict of a nerfstudio method config (dn_config.py:36-75) with the Gaussian groups' Adam configs
    replaced by FusedAdamOptimizerConfig (same lr / eps; schedulers untouched).  Returns the same dict."""
    for name in groups:
        if name in optimizers_config:
            old = optimizers_config[name]["optimizer"]
            optimizers_config[name]["optimizer"] = FusedAdamOptimizerConfig(
                lr=old.lr, eps=old.eps, max_norm=getattr(old, "max_norm", None),
                weight_decay=getattr(old, "weight_decay", 0))
    return optimizers_config
