"""Adaptive density control of the Gaussian set — host-side mirror of the reference model's
`after_train` / `refinement_after` (gs_toolkit/models/vanilla_gs.py:344-372, 381-497) on top of the device-side
compaction kernels of csrc/densify.cu.

The decisions (which Gaussians are split, duplicated, culled; the row order of the new set; the Adam moments of the
survivors; the opacity reset) are the reference's, including its quirks:
  * `split_gaussians` shrinks the scales of the split Gaussians IN PLACE before the duplication mask is evaluated
    (:567-569 vs :419-423), so a split Gaussian whose reduced scale is <= densify_size_thresh is duplicated as well;
  * new Gaussians enter the cull test with max_2Dsize = 0 (:433-441);
  * the running statistics are dropped after every refinement (:491-493).
The only host synchronisation is ONE 16-byte read (the four counters that size the new tensors); the reference
synchronises on `.sum().item()`, `torch.where` and every boolean-mask indexing.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass
from typing import Dict, Optional, Tuple

import torch
from torch import Tensor

from . import _lib
from .cuda import _Guard, _check_input, _ptr
from .optim import GaussianOptimizers

KIND_COPY, KIND_MEANS, KIND_SCALES, KIND_ZERO_NEW = 0, 1, 2, 3


@dataclass
class DensifyConfig:
    """The refinement fields of GaussianSplattingModelConfig, same names and defaults (vanilla_gs.py:44-87)."""

    warmup_length: int = 500
    refine_every: int = 100
    cull_alpha_thresh: float = 0.1
    cull_scale_thresh: float = 0.5
    continue_cull_post_densification: bool = True
    reset_alpha_every: int = 30
    densify_grad_thresh: float = 0.0002
    densify_size_thresh: float = 0.01
    n_split_samples: int = 2
    cull_screen_size: float = 0.15
    split_screen_size: float = 0.05
    stop_screen_size_at: int = 4000
    stop_split_at: int = 10_000


class DensifyStats:
    """xys_grad_norm / vis_counts / max_2Dsize of the reference model (None until the first view after a refinement)."""

    def __init__(self):
        self.xys_grad_norm: Optional[Tensor] = None
        self.vis_counts: Optional[Tensor] = None
        self.max_2Dsize: Optional[Tensor] = None

    def reset(self) -> None:
        self.xys_grad_norm = self.vis_counts = self.max_2Dsize = None

    def update(self, xys_grad: Tensor, radii: Tensor, last_size: Tuple[int, int]) -> None:
        """`after_train` (vanilla_gs.py:344-372).  xys_grad: [N,2] float32 — `xys.grad` of the drop-in operators, or
        `RenderAux.xys_grad` of the fused operator (a strided view into its gradient records is read in place);
        radii: [N] int32; last_size = (height, width) of the rendered view."""
        n = radii.numel()
        _check_input(radii, "radii", torch.int32)
        if not xys_grad.is_cuda or xys_grad.dtype != torch.float32 or xys_grad.dim() != 2 or xys_grad.shape != (n, 2) \
                or xys_grad.stride(1) != 1:
            raise RuntimeError(f"xys_grad must be a float32 CUDA tensor of shape ({n}, 2) with unit inner stride")
        first = self.xys_grad_norm is None
        if first:
            self.xys_grad_norm = torch.empty(n, dtype=torch.float32, device=radii.device)
            self.vis_counts = torch.empty_like(self.xys_grad_norm)
            self.max_2Dsize = torch.empty_like(self.xys_grad_norm)
        elif self.xys_grad_norm.numel() != n:
            raise RuntimeError("the Gaussian count changed without a refinement (statistics are stale)")
        with _Guard(radii) as st:
            _lib.check(_lib.load().gsr_densify_stats_update(
                n, _ptr(xys_grad), int(xys_grad.stride(0)), _ptr(radii), float(max(last_size[0], last_size[1])),
                int(first), _ptr(self.xys_grad_norm), _ptr(self.vis_counts), _ptr(self.max_2Dsize), st),
                "densify_stats_update")

    def all_reduce(self, group=None) -> None:
        """View-parallel training: make the statistics identical on every rank — sum, sum, max
        (view_parallel.all_reduce_densification_stats) — so that all replicas take the same split / cull decisions.
        Call it right before `refinement_after`; the ratio xys_grad_norm / vis_counts is then the mean gradient norm
        over all ranks' views."""
        from .view_parallel import all_reduce_densification_stats

        if self.xys_grad_norm is not None:
            all_reduce_densification_stats(self.xys_grad_norm, self.vis_counts, self.max_2Dsize, group)


def plan(params: Dict[str, Tensor], stats: DensifyStats, config: DensifyConfig, step: int, do_densify: bool,
         last_size: Tuple[int, int]):
    """Classification + scan: returns (flags [N] u8, ranks [N,4] i32, counts (n_split, n_keep_orig, n_keep_split,
    n_keep_dup) as python ints)."""
    scales, opac = params["scales"].detach(), params["opacities"].detach()
    _check_input(scales, "scales", torch.float32)
    _check_input(opac, "opacities", torch.float32)
    n, dev = scales.shape[0], scales.device
    use_screen = step < config.stop_screen_size_at
    cull_big = step > config.refine_every * config.reset_alpha_every
    if do_densify:
        assert stats.xys_grad_norm is not None and stats.vis_counts is not None and stats.max_2Dsize is not None
    if cull_big and use_screen:
        assert stats.max_2Dsize is not None
    lib = _lib.load()
    flags = torch.empty(n, dtype=torch.uint8, device=dev)
    ranks = torch.empty((n, 4), dtype=torch.int32, device=dev)
    counts = torch.empty(4, dtype=torch.int32, device=dev)
    ws_bytes = lib.gsr_densify_plan_workspace_bytes(n)
    ws = torch.empty(ws_bytes, dtype=torch.uint8, device=dev)
    null = C.c_void_p(0)
    gn = _ptr(stats.xys_grad_norm) if stats.xys_grad_norm is not None else null
    vc = _ptr(stats.vis_counts) if stats.vis_counts is not None else null
    m2 = _ptr(stats.max_2Dsize) if stats.max_2Dsize is not None else null
    with _Guard(scales) as st:
        _lib.check(lib.gsr_densify_plan(
            n, _ptr(scales), _ptr(opac), gn, vc, m2, int(do_densify), float(max(last_size[0], last_size[1])),
            config.densify_grad_thresh, config.densify_size_thresh, int(use_screen), config.split_screen_size,
            config.cull_alpha_thresh, int(cull_big), config.cull_scale_thresh,
            int(cull_big and use_screen and stats.max_2Dsize is not None), config.cull_screen_size, _ptr(flags),
            _ptr(ranks), _ptr(counts), _ptr(ws), ws_bytes, st), "densify_plan")
    c = counts.tolist()  # the one host read
    return flags, ranks, (int(c[0]), int(c[1]), int(c[2]), int(c[3]))


def apply(params: Dict[str, Tensor], optimizers: Optional[GaussianOptimizers], flags: Tensor, ranks: Tensor,
          counts: Tuple[int, int, int, int], n_split_samples: int, samples: Optional[Tensor]) -> int:
    """Gather the new Gaussian set and the new Adam moments; replaces the tensors of `params` (new leaves with
    requires_grad as before) and the moments held by `optimizers`.  Returns the new Gaussian count."""
    n = flags.numel()
    n_split, n_keep_orig, n_keep_split, n_keep_dup = counts
    new_n = n_keep_orig + n_split_samples * n_keep_split + n_keep_dup
    dev = flags.device
    if n_keep_split > 0:
        if samples is None or samples.shape != (n_split_samples * n_split, 3):
            raise ValueError(f"samples must have shape ({n_split_samples * n_split}, 3)")
        _check_input(samples, "samples", torch.float32)
    srcs, dsts, widths, kinds = [], [], [], []
    new_params: Dict[str, Tensor] = {}
    new_moments: Dict[str, Tuple[Tensor, Tensor]] = {}
    for name, p in params.items():
        src = p.detach()
        _check_input(src, name, torch.float32)
        if src.shape[0] != n:
            raise RuntimeError(f"{name}: expected {n} rows, got {src.shape[0]}")
        w = src.numel() // n if n else 1
        dst = torch.empty((new_n,) + tuple(src.shape[1:]), dtype=torch.float32, device=dev)
        new_params[name] = dst
        srcs.append(src)
        dsts.append(dst)
        widths.append(w)
        kinds.append(KIND_MEANS if name == "means" else KIND_SCALES if name == "scales" else KIND_COPY)
        if optimizers is not None and name in optimizers.state:
            m, v = optimizers.moments(name)
            if m is not None:
                nm, nv = torch.empty_like(dst), torch.empty_like(dst)
                new_moments[name] = (nm, nv)
                for s_, d_ in ((m, nm), (v, nv)):
                    srcs.append(s_)
                    dsts.append(d_)
                    widths.append(w)
                    kinds.append(KIND_ZERO_NEW)
    k = len(srcs)
    ptr_t, i32_t = C.c_void_p * k, C.c_int32 * k
    src_a, dst_a = ptr_t(*[t.data_ptr() for t in srcs]), ptr_t(*[t.data_ptr() for t in dsts])
    w_a, k_a = i32_t(*widths), i32_t(*kinds)
    counts_a = (C.c_int32 * 4)(*counts)
    fmap = torch.empty(max(new_n, 1), dtype=torch.int32, device=dev)
    null = C.c_void_p(0)
    with _Guard(flags) as st:
        _lib.check(_lib.load().gsr_densify_apply(
            n, n_split_samples, counts_a, _ptr(flags), _ptr(ranks), _ptr(samples) if samples is not None else null,
            _ptr(params["means"].detach()), _ptr(params["scales"].detach()), _ptr(params["quats"].detach()), k, src_a,
            dst_a, w_a, k_a, _ptr(fmap), st), "densify_apply")
    for name, p in list(params.items()):
        params[name] = new_params[name].requires_grad_(p.requires_grad)
    if optimizers is not None:
        for name, (nm, nv) in new_moments.items():
            optimizers.set_moments(name, nm, nv)
    return new_n


def refinement_after(params: Dict[str, Tensor], optimizers: Optional[GaussianOptimizers], stats: DensifyStats,
                     config: DensifyConfig, step: int, num_train_data: int, last_size: Tuple[int, int],
                     samples: Optional[Tensor] = None, generator: Optional[torch.Generator] = None) -> Dict[str, int]:
    """`GaussianSplattingModel.refinement_after` (vanilla_gs.py:381-497).  `params` is the model's parameter dict
    (means, scales, quats, features_dc, features_rest, opacities [+ extra per-Gaussian tensors, copied]); its tensors
    are replaced.  `samples` overrides the standard-normal draws of split_gaussians ([n_split_samples * n_split, 3]);
    by default they come from `torch.randn(..., device=..., generator=generator)` exactly as in the reference
    (:543-545), so the same torch seed gives the same new Gaussians.
    Returns counters {"n_before", "n_after", "n_split" (Gaussians split), "n_kept" (surviving originals),
    "n_new_split" / "n_new_dup" (surviving new rows), "opacity_reset"}."""
    n0 = int(params["means"].shape[0])
    info = {"n_before": n0, "n_after": n0, "n_split": 0, "n_kept": n0, "n_new_split": 0, "n_new_dup": 0,
            "opacity_reset": 0}
    if step <= config.warmup_length or n0 == 0:
        return info
    reset_interval = config.reset_alpha_every * config.refine_every
    do_densification = (step < config.stop_split_at
                        and step % reset_interval > num_train_data + config.refine_every)
    cull_only = (not do_densification) and step >= config.stop_split_at and config.continue_cull_post_densification
    if do_densification or cull_only:
        flags, ranks, counts = plan(params, stats, config, step, do_densification, last_size)
        n_split, n_keep_orig, n_keep_split, n_keep_dup = counts
        if do_densification and n_split > 0 and samples is None:
            samples = torch.randn((config.n_split_samples * n_split, 3), device=params["means"].device,
                                  generator=generator)
        if not (n_keep_orig == n0 and n_keep_split == 0 and n_keep_dup == 0):
            info["n_after"] = apply(params, optimizers, flags, ranks, counts, config.n_split_samples, samples)
        info.update(n_split=n_split, n_kept=n_keep_orig, n_new_split=config.n_split_samples * n_keep_split,
                    n_new_dup=n_keep_dup)
    if step < config.stop_split_at and step % reset_interval == config.refine_every:
        # :472-489 — reset value is twice the cull threshold; Adam moments of the opacities are zeroed
        reset_value = config.cull_alpha_thresh * 2.0
        max_logit = torch.logit(torch.tensor(reset_value, dtype=torch.float32)).item()
        op = params["opacities"].detach()
        m, v = optimizers.moments("opacities") if optimizers is not None else (None, None)
        null = C.c_void_p(0)
        with _Guard(op) as st:
            _lib.check(_lib.load().gsr_opacity_reset(op.numel(), float(max_logit), _ptr(op),
                                                     _ptr(m) if m is not None else null,
                                                     _ptr(v) if v is not None else null, st), "opacity_reset")
            torch.autograd.graph.increment_version(op)  # written through a raw pointer (see optim.py)
        info["opacity_reset"] = 1
    stats.reset()  # :491-493
    return info
