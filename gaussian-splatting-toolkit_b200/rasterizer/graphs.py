"""CUDA-graph capture of a training / rendering step built from this package's operators.

With the asynchronous binning mode (rasterizer.binning) a view — project_gaussians, spherical_harmonics, rasterize_gaussians
and their backward — neither reads the device from the host nor sizes a launch by a device-side count, so the whole step
can be captured once and replayed: one launch per view instead of ~40 Python-dispatched calls (the eager public path costs
0.8 - 2 ms of host time per 1080p view, about the GPU time of the view itself).  The reference cannot do this: it sizes its
pair buffers with `.item()` (rasterizer/utils.py:124).

    step = rasterizer.graphs.capture_step(fn)      # fn() uses only STATIC tensors (same storage every call)
    ... fill the static input tensors (e.g. asynchronous H2D copies) ...
    step.replay()                                  # outputs / .grad tensors are static too
    rasterizer.graphs.check()                      # once per iteration or less often: pair-buffer overflow of the replays

Streams: autograd ties a leaf's AccumulateGrad node to the stream on which the leaf is first used; if the parameters were
already used on another stream (e.g. eager iterations on the default stream), pass that stream's successor consistently
(`capture_step(fn, stream=s)` with the eager iterations under `torch.cuda.stream(s)`), or capture forward and backward
as two graphs as bench.py does.

Capacity: the pair buffers of a captured rasterize call have the capacity the signature (N, H, W, block_width) had learned
by the time of the capture (1.25 x the largest M seen + slack); a replay that needs more sets the overflow flag, which
`check()` turns into BinningOverflow (re-capture and repeat).  Re-capture after densification (N changes).
"""
from __future__ import annotations

from typing import Callable, Optional

import torch

from . import binning


class CapturedStep:
    """A captured step: `.replay()` runs it; `.graph` is the torch.cuda.CUDAGraph; `.result` is what `fn` returned during
    capture (static tensors)."""

    def __init__(self, graph, result):
        self.graph, self.result = graph, result

    def replay(self):
        self.graph.replay()
        return self.result


def capture_step(fn: Callable[[], object], warmup: int = 3, pool=None, stream: Optional["torch.cuda.Stream"] = None) -> CapturedStep:
    """Run `fn` `warmup` times eagerly on a side stream (allocator warm-up; the asynchronous binning learns the pair-buffer
    capacity of every rasterize signature `fn` uses), then capture one more call into a CUDA graph.  Switches the package to
    the asynchronous binning mode (a captured step cannot read M on the host)."""
    binning.set_binning_mode("async")
    binning.prepare_capture_slots()
    side = stream if stream is not None else torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        for _ in range(max(2, warmup)):  # the first call of a signature is synchronous; the second uses the learned capacity
            fn()
    torch.cuda.current_stream().wait_stream(side)
    torch.cuda.synchronize()
    binning.check()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g, pool=pool, stream=side):
        result = fn()
    torch.cuda.current_stream().wait_stream(side)
    return CapturedStep(g, result)


def check(synchronize: bool = True) -> None:
    """Pair-buffer overflow check of the replayed graphs (see rasterizer.binning.check_captured)."""
    binning.check_captured(synchronize)
