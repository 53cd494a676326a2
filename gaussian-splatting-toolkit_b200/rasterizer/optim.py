"""`GaussianOptimizers` — the optimizer step of the Gaussian parameter groups as ONE kernel.

Host-side mirror of `gs_toolkit.engine.optimizers.Optimizers` (engine/optimizers.py:59-214) restricted to what the
Gaussian models use: one Adam "optimizer" per parameter group (`means`, `features_dc`, `features_rest`, `opacities`,
`scales`, `quats`; learning rates and eps = 1e-15 from configs/method_configs.py:98-125), an optional
`ExponentialDecayScheduler` per group (engine/schedulers.py:94-135; the reference puts one on `means`), and the
state surgery densification needs.  `optimizer_step_all()` issues a single `gsr_adam_step_multi` launch over all
groups instead of torch's ~70 foreach kernels; the arithmetic is torch.optim.Adam's.

State layout is torch.optim.Adam's (`exp_avg`, `exp_avg_sq`, `step`), and `state_dict()` / `load_state_dict()` speak
the per-group format the reference checkpoints hold (`{"optimizers": {group: adam.state_dict()}}`,
engine/trainer.py:459-469), so a checkpoint written by either side loads on the other.
"""
from __future__ import annotations

import ctypes as C
import math
from typing import Callable, Dict, Optional

import numpy as np
import torch
from torch import Tensor

from . import _lib
from .cuda import _Guard, _check_input, _ptr

# configs/method_configs.py:98-125
DEFAULT_LRS = {"means": 1.6e-4, "features_dc": 0.0025, "features_rest": 0.0025 / 20, "opacities": 0.05,
               "scales": 0.005, "quats": 0.001}
GROUPS = ("means", "scales", "quats", "features_dc", "features_rest", "opacities")  # vanilla_gs.py:623-636


def exponential_decay_lr(lr_init: float, lr_final: Optional[float], max_steps: int, warmup_steps: int = 0,
                         lr_pre_warmup: float = 1e-8, ramp: str = "cosine") -> Callable[[int], float]:
    """Learning rate at scheduler step `step` of ExponentialDecayScheduler (engine/schedulers.py:101-135); the
    reference multiplies lr_init by (this / lr_init) through LambdaLR."""
    lr_final = lr_init if lr_final is None else lr_final

    def lr_at(step: int) -> float:
        if step < warmup_steps:
            if ramp == "cosine":
                return lr_pre_warmup + (lr_init - lr_pre_warmup) * math.sin(
                    0.5 * math.pi * min(max(step / warmup_steps, 0.0), 1.0))
            return lr_pre_warmup + (lr_init - lr_pre_warmup) * step / warmup_steps
        t = min(max((step - warmup_steps) / (max_steps - warmup_steps), 0.0), 1.0)
        lr = float(np.exp(np.log(lr_init) * (1 - t) + np.log(lr_final) * t))
        return lr_init * (lr / lr_init)  # LambdaLR: initial_lr * multiplier (schedulers.py:130-134)

    return lr_at


def default_means_scheduler() -> Callable[[int], float]:
    """configs/method_configs.py:100-105: means decay 1.6e-4 -> 1.6e-6 over 30000 steps."""
    return exponential_decay_lr(1.6e-4, 1.6e-6, 30000)


class GaussianOptimizers:
    """One Adam state per named parameter group, stepped together by one kernel.

    `params` maps group name -> float32 CUDA tensor (leaf, requires_grad); the dict is kept by reference so that
    `rasterizer.densify.refinement_after` can replace the tensors (as the reference replaces the Parameters,
    vanilla_gs.py:424-431, 523-524)."""

    def __init__(self, params: Dict[str, Tensor], lrs: Optional[Dict[str, float]] = None, eps: float = 1e-15,
                 betas=(0.9, 0.999), schedulers: Optional[Dict[str, Callable[[int], float]]] = None):
        lrs = dict(DEFAULT_LRS) if lrs is None else dict(lrs)
        for name in params:
            if name not in lrs:
                # engine/optimizers.py:88-91
                raise RuntimeError(f"Optimizer config for '{name}' not found in config file. Make sure you specify an "
                                   f"optimizer for each parameter group. Provided configs were: {lrs.keys()}")
        if len(params) > 8:
            raise RuntimeError("at most 8 parameter groups (GSR_ADAM_MAX_SEGMENTS)")
        self.params = params
        self.lr_init = {k: float(lrs[k]) for k in params}
        self.lrs = dict(self.lr_init)
        self.eps, self.betas = float(eps), (float(betas[0]), float(betas[1]))
        self.schedulers = dict(schedulers or {})
        self.sched_steps = {k: 0 for k in self.schedulers}
        self.state: Dict[str, Dict[str, object]] = {}
        for name, p in params.items():
            _check_input(p.detach(), name, torch.float32)
            self.state[name] = {"step": 0, "exp_avg": None, "exp_avg_sq": None}  # created lazily, like torch

    # ------------------------------------------------------------------ stepping
    def optimizer_step_all(self, grads: Optional[Dict[str, Tensor]] = None, grad_scale: float = 1.0) -> None:
        """engine/optimizers.py:173-180.  `grads` defaults to the `.grad` of every parameter; groups without a
        gradient are skipped (torch.optim skips parameters whose grad is None).  `grad_scale` multiplies every
        gradient (1 / world_size after a summing all-reduce)."""
        names = []
        for name, p in self.params.items():
            g = grads[name] if grads is not None else p.grad
            if g is None:
                continue
            names.append((name, p, g))
        if not names:
            return
        n = len(names)
        ptr_t, i64_t, f64_t = C.c_void_p * n, C.c_int64 * n, C.c_double * n
        ps, gs, ms, vs, numels, lrs, steps = ptr_t(), ptr_t(), ptr_t(), ptr_t(), i64_t(), f64_t(), i64_t()
        keep = []
        for k, (name, p, g) in enumerate(names):
            st = self.state[name]
            if st["exp_avg"] is None:
                st["exp_avg"] = torch.zeros_like(p, memory_format=torch.contiguous_format)
                st["exp_avg_sq"] = torch.zeros_like(p, memory_format=torch.contiguous_format)
            g = g.detach()
            if g.dtype != torch.float32 or not g.is_contiguous():
                g = g.float().contiguous()
            if g.shape != p.shape:
                raise RuntimeError(f"{name}: gradient shape {tuple(g.shape)} != parameter shape {tuple(p.shape)}")
            _check_input(p.detach(), name, torch.float32)
            keep.append(g)
            st["step"] = int(st["step"]) + 1
            ps[k], gs[k], ms[k], vs[k] = p.data_ptr(), g.data_ptr(), st["exp_avg"].data_ptr(), st["exp_avg_sq"].data_ptr()
            numels[k], lrs[k], steps[k] = p.numel(), self.lrs[name], st["step"]
        with _Guard(names[0][1]) as stream:
            _lib.check(_lib.load().gsr_adam_step_multi(n, ps, gs, ms, vs, numels, lrs, steps, self.betas[0],
                                                       self.betas[1], self.eps, float(grad_scale), stream),
                       "adam_step_multi")
        # the kernel wrote through raw pointers: bump the autograd version counters, so that anything keyed on them (the
        # binning cache of rasterize.py, autograd's saved-tensor check) sees the parameters as modified
        for name, p, _ in names:
            torch.autograd.graph.increment_version(p)

    step = optimizer_step_all

    def zero_grad_all(self) -> None:
        """engine/optimizers.py:117-120 (set_to_none semantics of torch >= 2.0)."""
        for p in self.params.values():
            p.grad = None

    def scheduler_step_all(self, step: int = 0) -> None:
        """engine/optimizers.py:182-195: advance every scheduler by one and set the group's learning rate."""
        for name, fn in self.schedulers.items():
            self.sched_steps[name] += 1
            self.lrs[name] = fn(self.sched_steps[name])

    def scheduler_state_dict(self) -> Dict[str, dict]:
        """{group: LambdaLR.state_dict()} as saved by engine/trainer.py:470 (`{k: v.state_dict() for k, v in
        self.optimizers.schedulers.items()}`): `last_epoch` carries the number of scheduler steps taken."""
        return {name: {"base_lrs": [self.lr_init[name]], "last_epoch": int(self.sched_steps[name]),
                       "_step_count": int(self.sched_steps[name]) + 1, "_get_lr_called_within_step": False,
                       "_last_lr": [self.lrs[name]], "lr_lambdas": [None]} for name in self.schedulers}

    def load_schedulers(self, loaded: Dict[str, dict]) -> None:
        """engine/optimizers.py:207-214: restore the scheduler progress (and the learning rate it had produced), so that
        the next scheduler_step_all() continues the decay instead of restarting it."""
        for name, sd in loaded.items():
            if name not in self.schedulers:
                continue
            self.sched_steps[name] = int(sd["last_epoch"])
            last = sd.get("_last_lr")
            self.lrs[name] = float(last[0]) if last else float(self.schedulers[name](max(self.sched_steps[name], 0)))

    # ------------------------------------------------------------------ densification support
    def moments(self, name: str):
        st = self.state[name]
        return st["exp_avg"], st["exp_avg_sq"]

    def set_moments(self, name: str, exp_avg: Optional[Tensor], exp_avg_sq: Optional[Tensor]) -> None:
        self.state[name]["exp_avg"], self.state[name]["exp_avg_sq"] = exp_avg, exp_avg_sq

    # ------------------------------------------------------------------ checkpoint format
    def state_dict(self) -> Dict[str, dict]:
        """{group: torch.optim.Adam.state_dict()} as saved by engine/trainer.py:459-469."""
        out = {}
        for name in self.params:
            st = self.state[name]
            state = {}
            if st["exp_avg"] is not None:
                state[0] = {"step": torch.tensor(float(st["step"])), "exp_avg": st["exp_avg"], "exp_avg_sq": st["exp_avg_sq"]}
            out[name] = {"state": state, "param_groups": [{
                "lr": self.lrs[name], "betas": self.betas, "eps": self.eps, "weight_decay": 0, "amsgrad": False,
                "maximize": False, "foreach": None, "capturable": False, "differentiable": False, "fused": None,
                "initial_lr": self.lr_init[name], "params": [0]}]}
        return out

    def load_state_dict(self, loaded: Dict[str, dict]) -> None:
        """engine/optimizers.py:197-204."""
        for name, sd in loaded.items():
            if name not in self.params:
                continue
            p = self.params[name]
            st = sd["state"].get(0)
            if st is not None:
                self.state[name] = {"step": int(float(st["step"])),
                                    "exp_avg": st["exp_avg"].to(p.device, torch.float32).contiguous().clone(),
                                    "exp_avg_sq": st["exp_avg_sq"].to(p.device, torch.float32).contiguous().clone()}
            else:
                self.state[name] = {"step": 0, "exp_avg": None, "exp_avg_sq": None}
            self.lrs[name] = float(sd["param_groups"][0]["lr"])
