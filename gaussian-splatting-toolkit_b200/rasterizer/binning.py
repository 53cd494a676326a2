"""Internal tile binning of `rasterize_gaussians` / `render_gaussians`: synchronous (exact buffers) or asynchronous
(device-side pair count, capacity-bounded buffers).

The reference learns the number of (Gaussian, tile) pairs M with `.item()` (rasterizer/utils.py:124) — one stream
synchronisation per rasterize call — because it must size the pair buffers.  Two modes here:

``sync`` (default)  the same contract as the reference: one host read of M per call, buffers of exactly M entries.
``async``           opt-in (`set_binning_mode("async")` or GSR_BINNING=async): no host read.  The pair buffers have a
                    CAPACITY learned from earlier calls with the same (N, H, W, block_width) signature (the first call
                    of a signature runs synchronously); every call leaves {M, overflow} in a pinned host slot through
                    an asynchronous copy, and the slots of earlier calls are inspected — without blocking — at the
                    start of the next call, in `poll()` and in the backward pass:
                      * M above 80 % of the capacity  -> the capacity grows for the following calls;
                      * overflow (M > capacity: the farthest pairs of THAT call were dropped) -> `BinningOverflow` is
                        raised, the capacity is raised to 1.5 M, and the caller repeats the iteration.
                    A training loop calls `rasterizer.binning.poll()` (non-blocking) or `check()` (blocking) once per
                    iteration before the optimizer step; the whole view is then free of host synchronisation and can be
                    captured in a CUDA graph.
"""
from __future__ import annotations

import os
from typing import Dict, List, Tuple

import torch

from . import cuda as _C

_MODE = "async" if os.environ.get("GSR_BINNING", "sync").lower() == "async" else "sync"
HEADROOM = 1.25       # capacity = HEADROOM * largest M seen (+ a small constant)
GROW_AT = 0.80        # grow when a call used more than this fraction of its capacity


class BinningOverflow(RuntimeError):
    """An earlier asynchronous rasterize call had more (Gaussian, tile) pairs than its buffers could hold."""


def set_binning_mode(mode: str) -> None:
    global _MODE
    if mode not in ("sync", "async"):
        raise ValueError("binning mode must be 'sync' or 'async'")
    _MODE = mode


def get_binning_mode() -> str:
    return _MODE


class _Signature:
    __slots__ = ("capacity", "max_seen")

    def __init__(self):
        self.capacity = 0
        self.max_seen = 0


_signatures: Dict[Tuple, _Signature] = {}
_pending: List[Tuple[torch.Tensor, "torch.cuda.Event", Tuple, int]] = []  # (pinned meta, event, signature key, capacity)
_free_slots: List[torch.Tensor] = []
_RING = 64  # pinned {M, overflow, ..} slots, allocated ONCE (pinning is a slow, synchronising call): the host can run
            # at most _RING asynchronous rasterize calls ahead of the device


def _slot() -> torch.Tensor:
    global _free_slots
    if not _free_slots and not _pending:
        ring = torch.zeros(_RING, 4, dtype=torch.int32).pin_memory()
        _free_slots = [ring[i] for i in range(_RING)]
    if not _free_slots:
        _inspect(block=False)
    if not _free_slots:  # every slot belongs to a call still in flight: wait for the oldest one
        _pending[0][1].synchronize()
        _inspect(block=False)
    return _free_slots.pop()


def _inspect(block: bool) -> None:
    """Look at the finished calls (all of them when `block`), adapt capacities, raise on overflow."""
    overflow = None
    keep = []
    for meta, ev, key, cap in _pending:
        if not block and not ev.query():
            keep.append((meta, ev, key, cap))
            continue
        if block:
            ev.synchronize()
        m, over = int(meta[0]), int(meta[1])
        _free_slots.append(meta)
        sig = _signatures.get(key)
        if sig is not None:
            sig.max_seen = max(sig.max_seen, m)
            if m > GROW_AT * sig.capacity:
                sig.capacity = max(sig.capacity, int(HEADROOM * m) + 65536)
            if over:
                sig.capacity = max(sig.capacity, int(1.5 * m) + 65536)
        if over and overflow is None:
            overflow = (m, cap)
    _pending[:] = keep
    if overflow is not None:
        raise BinningOverflow(
            f"an asynchronous rasterize call produced {overflow[0]} (Gaussian, tile) pairs but its buffers held "
            f"{overflow[1]}: its image / gradients were computed from a truncated list.  The capacity has been raised; "
            f"repeat the iteration (or use set_binning_mode('sync')).")


_capture_slots: List[Tuple[torch.Tensor, Tuple, int]] = []  # pinned {M, overflow} slots written by graph replays
_capture_slots_free: List[torch.Tensor] = []


def prepare_capture_slots(n: int = 8) -> None:
    """Pin `n` slots BEFORE a capture starts (pinning is a synchronising allocation and must not happen inside one)."""
    while len(_capture_slots_free) < n:
        _capture_slots_free.append(torch.zeros(4, dtype=torch.int32).pin_memory())


def check_captured(synchronize: bool = True) -> None:
    """After replaying graphs captured over rasterize_gaussians: raise BinningOverflow if the LAST replay of any of them
    produced more (Gaussian, tile) pairs than the capacity baked into the graph (the capacity of the signature is raised;
    re-capture and repeat).  Reads pinned memory written by the replays, so it synchronises the device first."""
    if synchronize and _capture_slots:
        torch.cuda.synchronize()
    for slot, key, cap in _capture_slots:
        m, over = int(slot[0]), int(slot[1])
        sig = _signatures.get(key)
        if sig is not None:
            sig.max_seen = max(sig.max_seen, m)
            if over:
                sig.capacity = max(sig.capacity, int(1.5 * m) + 65536)
        if over:
            raise BinningOverflow(f"a replayed CUDA graph produced {m} (Gaussian, tile) pairs but its buffers hold {cap}: "
                                  f"re-capture the step (the capacity has been raised) and repeat the iteration")


def poll() -> None:
    """Non-blocking: inspect the asynchronous calls that have finished; raises BinningOverflow if one overflowed."""
    if torch.cuda.is_current_stream_capturing():
        return
    _inspect(block=False)


def check() -> None:
    """Blocking: wait for every outstanding asynchronous call and inspect it."""
    _inspect(block=True)


def reset() -> None:
    """Forget learned capacities and outstanding tickets (tests)."""
    _signatures.clear()
    _pending.clear()
    _capture_slots.clear()


def bin_gaussians(xys, depths, radii, conics, opacity, img_height, img_width, block_width):
    """-> (num_intersects or None, gaussian_ids_sorted, tile_bins).  `num_intersects` is an int in sync mode (0 => the
    other two are None) and None in async mode (unknown on the host; empty scenes simply have empty tile bins)."""
    op = opacity.reshape(-1)
    if _MODE == "sync":
        return _C.bin_gaussians_fast(xys, depths, radii, conics, op, img_height, img_width, block_width)
    dev = xys.device
    key = (dev.index, xys.size(0), img_height, img_width, block_width)
    if torch.cuda.is_current_stream_capturing():
        # CUDA-graph capture (rasterizer.graphs): no host-side bookkeeping may run — the capacity must have been learned by
        # eager warm-up calls, and {M, overflow} of every REPLAY lands in a pinned slot owned by the capture
        sig = _signatures.get(key)
        if sig is None:
            raise RuntimeError("rasterize_gaussians under CUDA-graph capture: run the step eagerly first (asynchronous "
                               "binning learns its pair-buffer capacity from the first call of a signature)")
        if not _capture_slots_free:
            raise RuntimeError("rasterize_gaussians under CUDA-graph capture: call rasterizer.binning.prepare_capture_slots() "
                               "before the capture (pinning host memory inside a capture is not allowed)")
        slot = _capture_slots_free.pop()
        _capture_slots.append((slot, key, sig.capacity))
        ids, bins, _meta = _C.bin_gaussians_device(xys, depths, radii, conics, op, img_height, img_width, block_width,
                                                   sig.capacity, meta_pinned=slot)
        return None, ids, bins
    _inspect(block=False)
    sig = _signatures.get(key)
    if sig is None:
        # first call of this signature: learn M synchronously (exact buffers), size the capacity from it
        m, ids, bins = _C.bin_gaussians_fast(xys, depths, radii, conics, op, img_height, img_width, block_width)
        sig = _signatures[key] = _Signature()
        sig.max_seen = m
        sig.capacity = int(HEADROOM * m) + 65536
        return m, ids, bins
    meta_host = _slot()
    ids, bins, _meta = _C.bin_gaussians_device(xys, depths, radii, conics, op, img_height, img_width, block_width,
                                               sig.capacity, meta_pinned=meta_host)
    ev = torch.cuda.Event()
    ev.record(torch.cuda.current_stream(dev))
    _pending.append((meta_host, ev, key, sig.capacity))
    return None, ids, bins
