"""View-parallel multi-GPU plumbing: every rank holds the full, replicated Gaussian set and
renders its own training view; the only exchange step is ONE all-reduce (sum) of the parameter gradients —
59 floats per Gaussian (SH 48 + means 3 + scales 3 + quats 4 + opacity 1 at SH degree 3) — after backward.

Reference precedent: torch DDP's bucketed all-reduce of the `gauss_params` gradients
(gs_toolkit/pipelines/base_pipeline.py:202-207, implicit in loss.backward(), engine/trainer.py:498); views
are drawn per rank because `random` is seeded with seed + global_rank (scripts/train.py:54).  DDP averages;
`GradientBucket.all_reduce(average=True)` does the same.

Here the gradients live in ONE flat FP32 buffer whose segments are handed to the backward kernels as their
output pointers (the C ABI takes raw output pointers), so there is no per-tensor bucketing or copy: one
`ncclAllReduce` over the whole buffer, or two (SH segment first, on a side stream, overlapping the projection
backward) with `overlap=True`.  Works on any torch.distributed backend (NCCL on GPUs, gloo in the CPU tests).
"""
from __future__ import annotations

from typing import Dict, Optional, Tuple

import torch
import torch.distributed as dist

# segment order: the SH gradient (81 % of the bytes) first so that its all-reduce can start first
SEGMENTS = ("v_coeffs", "v_mean3d", "v_scale", "v_quat", "v_opacity")


def segment_shapes(num_points: int, sh_bases: int) -> Dict[str, Tuple[int, ...]]:
    return {"v_coeffs": (num_points, sh_bases, 3), "v_mean3d": (num_points, 3), "v_scale": (num_points, 3),
            "v_quat": (num_points, 4), "v_opacity": (num_points, 1)}


def floats_per_gaussian(sh_bases: int) -> int:
    return 3 * sh_bases + 3 + 3 + 4 + 1


class GradientBucket:
    """Flat gradient buffer + typed views into it."""

    def __init__(self, num_points: int, sh_bases: int = 16, device="cuda"):
        self.num_points, self.sh_bases = num_points, sh_bases
        self.flat = torch.zeros(floats_per_gaussian(sh_bases) * num_points, dtype=torch.float32, device=device)
        self.views: Dict[str, torch.Tensor] = {}
        self.offsets: Dict[str, Tuple[int, int]] = {}
        off = 0
        for name, shape in segment_shapes(num_points, sh_bases).items():
            n = 1
            for d in shape:
                n *= d
            # every segment starts on a 16-byte boundary (n is a multiple of 4 floats whenever N is)
            self.views[name] = self.flat[off:off + n].view(shape)
            self.offsets[name] = (off, off + n)
            off += n
        assert off == self.flat.numel()

    def __getitem__(self, name: str) -> torch.Tensor:
        return self.views[name]

    def pack(self, grads: Dict[str, torch.Tensor]) -> None:
        """Copy gradients produced elsewhere into the bucket (no-op for tensors that already alias it)."""
        for name in SEGMENTS:
            g = grads[name]
            v = self.views[name]
            if g.data_ptr() != v.data_ptr():
                v.copy_(g.reshape(v.shape))

    def all_reduce(self, group=None, average: bool = False, overlap_stream: Optional["torch.cuda.Stream"] = None):
        """Sum (or average) the bucket over the ranks.  With `overlap_stream` the SH segment is reduced on that
        stream first (call after the SH backward has been enqueued, before the projection backward) and the
        rest on the current stream; returns when both are enqueued."""
        if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
            return
        world = dist.get_world_size(group)
        if overlap_stream is None:
            dist.all_reduce(self.flat, group=group)
        else:
            lo, hi = self.offsets["v_coeffs"]
            cur = torch.cuda.current_stream()
            overlap_stream.wait_stream(cur)
            with torch.cuda.stream(overlap_stream):
                dist.all_reduce(self.flat[lo:hi], group=group)
            dist.all_reduce(self.flat[hi:], group=group)
            cur.wait_stream(overlap_stream)
        if average:
            self.flat.div_(world)


def view_for_rank(step: int, rank: int, world_size: int, num_views: int) -> int:
    """Round-robin view assignment: rank r renders view (step * world + r) mod num_views."""
    return (step * world_size + rank) % num_views


def all_reduce_densification_stats(xys_grad_norm: torch.Tensor, vis_counts: torch.Tensor,
                                   max_2dsize: torch.Tensor, group=None) -> None:
    """Replicas must take identical split / cull decisions: sum, sum, max over ranks of the statistics the
    reference model keeps per Gaussian (gs_toolkit/models/vanilla_gs.py:344-372)."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return
    dist.all_reduce(xys_grad_norm, group=group)
    dist.all_reduce(vis_counts, group=group)
    dist.all_reduce(max_2dsize, op=dist.ReduceOp.MAX, group=group)
