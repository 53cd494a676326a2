"""Tile binning utilities — same surface as the reference `rasterizer.utils` (rasterizer/utils.py:12-182)."""
from __future__ import annotations

from typing import Tuple

import torch
from torch import Tensor

from . import cuda as _C


def map_gaussian_to_intersects(num_points: int, num_intersects: int, xys: Tensor, depths: Tensor, radii: Tensor,
                               cum_tiles_hit: Tensor, tile_bounds: Tuple[int, int, int], block_size: int
                               ) -> Tuple[Tensor, Tensor]:
    """(tile << 32 | depth bits) keys [M] int64 and Gaussian ids [M] int32 (rasterizer/utils.py:12-52)."""
    return _C.map_gaussian_to_intersects(num_points, num_intersects, xys.contiguous(), depths.contiguous(),
                                         radii.contiguous(), cum_tiles_hit.contiguous(), tile_bounds, block_size)


def get_tile_bin_edges(num_intersects: int, isect_ids_sorted: Tensor, tile_bounds: Tuple[int, int, int]) -> Tensor:
    """tile_bins [T,2] int32: [first, last+1) of every tile in the sorted list (rasterizer/utils.py:55-81)."""
    return _C.get_tile_bin_edges(num_intersects, isect_ids_sorted.contiguous(), tile_bounds)


def compute_cov2d_bounds(cov2d: Tensor) -> Tuple[Tensor, Tensor]:
    """cov2d [N,3] (upper triangular) -> (conics [N,3], radii [N,1]) (rasterizer/utils.py:84-103)."""
    assert cov2d.shape[-1] == 3, (
        f"Expected input cov2d to be of shape (*batch, 3) (upper triangular values), but got {tuple(cov2d.shape)}")
    num_pts = cov2d.shape[0]
    assert num_pts > 0
    return _C.compute_cov2d_bounds(num_pts, cov2d.contiguous())


# one pinned int32 per device for the (unavoidable, API-mandated) host read of the intersection count
_pinned_total = {}


def compute_cumulative_intersects(num_tiles_hit: Tensor) -> Tuple[int, Tensor]:
    """(M, cum_tiles_hit) — int32 inclusive scan + host read of the total (rasterizer/utils.py:106-125).

    The scan is libgsr_b200's own (csrc/radix_sort.cuh); the total travels through a pinned host word filled by an async
    copy on the same stream, followed by a stream (not device) synchronisation."""
    nth = num_tiles_hit.contiguous()
    if nth.numel() == 0:
        return 0, torch.empty_like(nth, dtype=torch.int32)
    if nth.dtype != torch.int32:
        nth = nth.to(torch.int32)
    dev = nth.device
    key = (dev.index, torch.cuda.current_stream(dev).cuda_stream)
    pin = _pinned_total.get(key)
    if pin is None:
        pin = _pinned_total[key] = torch.zeros(1, dtype=torch.int32).pin_memory()
    cum = _C.cumsum_tiles_hit(nth, pin)
    torch.cuda.current_stream(dev).synchronize()
    return int(pin.item()), cum


def bin_and_sort_gaussians(num_points: int, num_intersects: int, xys: Tensor, depths: Tensor, radii: Tensor,
                           cum_tiles_hit: Tensor, tile_bounds: Tuple[int, int, int], block_size: int
                           ) -> Tuple[Tensor, Tensor, Tensor, Tensor, Tensor]:
    """(isect_ids, gaussian_ids, isect_ids_sorted, gaussian_ids_sorted, tile_bins) (rasterizer/utils.py:128-182).

    The reference sorts with torch.sort (returning an int64 permutation) and gathers the ids through it; here
    the (key, id) pairs go through one stable radix sort restricted to the bits that can be set."""
    isect_ids, gaussian_ids = map_gaussian_to_intersects(num_points, num_intersects, xys, depths, radii,
                                                         cum_tiles_hit, tile_bounds, block_size)
    num_tiles = int(tile_bounds[0]) * int(tile_bounds[1])
    isect_ids_sorted, gaussian_ids_sorted = _C.sort_intersects(isect_ids, gaussian_ids, num_tiles)
    tile_bins = get_tile_bin_edges(num_intersects, isect_ids_sorted, tile_bounds)
    return isect_ids, gaussian_ids, isect_ids_sorted, gaussian_ids_sorted, tile_bins
