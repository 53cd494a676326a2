"""`rasterizer` — B200-native drop-in for the package of the same name that Gaussian-Splatting-Toolkit's
models import (reference: gs_toolkit/gs_components/rasterizer/__init__.py:17-38).

Public surface (identical names, argument order and return values):
    project_gaussians, rasterize_gaussians, spherical_harmonics,
    map_gaussian_to_intersects, bin_and_sort_gaussians, compute_cumulative_intersects,
    compute_cov2d_bounds, get_tile_bin_edges, __version__,
    and the deprecated torch.autograd.Function shells of rasterizer/__init__.py:43-166.
Sub-modules the models import directly: rasterizer.sh (num_sh_bases, spherical_harmonics),
rasterizer.project_gaussians, rasterizer.rasterize, rasterizer._torch_impl (quat_to_rotmat).

Beyond the drop-in surface (not part of the reference package; a model opts in):
rasterizer.fused (one-node render operator), rasterizer.losses (fused L1 + SSIM), rasterizer.optim (one-launch Adam for
the parameter groups), rasterizer.densify (after_train / refinement_after as compaction kernels), rasterizer.io_ply /
rasterizer.io_scene (.ply, checkpoints, transforms.json), rasterizer.view_parallel (multi-GPU gradient exchange),
rasterizer.binning (`set_binning_mode("async")`: rasterize without the host read of the pair count).

All compute happens in libgsr_b200.so (hand-written sm_100a CUDA, C ABI in include/gsr_b200.h); importing
this package does not load the library (so CPU-only tooling can import it) but the first operator call
does, and raises if it is missing — there is no PyTorch or CPU fallback.
"""
from __future__ import annotations

import warnings
from typing import Any

import torch

from .binning import BinningOverflow, get_binning_mode, set_binning_mode
from .gaussian_rasterizer import GaussianRasterizationSettings, GaussianRasterizer
from .project_gaussians import project_gaussians
from .rasterize import rasterize_gaussians
from .sh import spherical_harmonics
from .utils import (
    bin_and_sort_gaussians,
    compute_cov2d_bounds,
    compute_cumulative_intersects,
    get_tile_bin_edges,
    map_gaussian_to_intersects,
)
from .version import __version__

__all__ = [
    "__version__",
    "project_gaussians",
    "rasterize_gaussians",
    "spherical_harmonics",
    "bin_and_sort_gaussians",
    "compute_cumulative_intersects",
    "compute_cov2d_bounds",
    "get_tile_bin_edges",
    "map_gaussian_to_intersects",
    "ProjectGaussians",
    "RasterizeGaussians",
    "BinAndSortGaussians",
    "ComputeCumulativeIntersects",
    "ComputeCov2dBounds",
    "GetTileBinEdges",
    "MapGaussiansToIntersects",
    "SphericalHarmonics",
    "NDRasterizeGaussians",
    # Inria-style façade (not part of the reference package; see gaussian_rasterizer.py)
    "GaussianRasterizer",
    "GaussianRasterizationSettings",
    # asynchronous tile binning (no host read of the pair count; see binning.py)
    "set_binning_mode",
    "get_binning_mode",
    "BinningOverflow",
]


def _deprecated_shell(cls_name: str, fn, fn_name: str):
    """Backwards-compatible `X.apply(...)` shell that warns and forwards (rasterizer/__init__.py:43-166)."""

    def forward(ctx, *args, **kwargs):
        warnings.warn(f"{cls_name} is deprecated, use {fn_name} instead", DeprecationWarning)
        return fn(*args, **kwargs)

    def backward(ctx: Any, *grad_outputs: Any) -> Any:
        raise NotImplementedError

    return type(cls_name, (torch.autograd.Function,),
                {"forward": staticmethod(forward), "backward": staticmethod(backward), "__doc__": forward.__doc__})


MapGaussiansToIntersects = _deprecated_shell("MapGaussiansToIntersects", map_gaussian_to_intersects,
                                             "map_gaussian_to_intersects")
ComputeCumulativeIntersects = _deprecated_shell("ComputeCumulativeIntersects", compute_cumulative_intersects,
                                                "compute_cumulative_intersects")
ComputeCov2dBounds = _deprecated_shell("ComputeCov2dBounds", compute_cov2d_bounds, "compute_cov2d_bounds")
GetTileBinEdges = _deprecated_shell("GetTileBinEdges", get_tile_bin_edges, "get_tile_bin_edges")
BinAndSortGaussians = _deprecated_shell("BinAndSortGaussians", bin_and_sort_gaussians, "bin_and_sort_gaussians")
ProjectGaussians = _deprecated_shell("ProjectGaussians", project_gaussians, "project_gaussians")
RasterizeGaussians = _deprecated_shell("RasterizeGaussians", rasterize_gaussians, "rasterize_gaussians")
NDRasterizeGaussians = _deprecated_shell("NDRasterizeGaussians", rasterize_gaussians, "rasterize_gaussians")
SphericalHarmonics = _deprecated_shell("SphericalHarmonics", spherical_harmonics, "spherical_harmonics")
