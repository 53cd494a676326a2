"""Spherical-harmonics colours — same surface as the reference `rasterizer.sh` (rasterizer/sh.py:10-97)."""
from __future__ import annotations

from torch import Tensor
from torch.autograd import Function

from . import cuda as _C

_BASES = {0: 1, 1: 4, 2: 9, 3: 16}
_DEGREES = {1: 0, 4: 1, 9: 2, 16: 3, 25: 4}


def num_sh_bases(degree: int) -> int:
    """Number of SH bases of `degree` (25 for anything above 3, like rasterizer/sh.py:10-19)."""
    return _BASES.get(degree, 25)


def deg_from_sh(num_bases: int) -> int:
    """Inverse of num_sh_bases (rasterizer/sh.py:22-33)."""
    assert num_bases in _DEGREES, "Invalid number of SH bases"
    return _DEGREES[num_bases]


def spherical_harmonics(degrees_to_use: int, viewdirs: Tensor, coeffs: Tensor) -> Tensor:
    """Evaluate SH colours.  viewdirs [N,3] (normalised inside the kernel), coeffs [N,K,3] -> [N,3].

    Differentiable w.r.t. `coeffs` only (rasterizer/sh.py:36-59)."""
    assert coeffs.shape[-2] >= num_sh_bases(degrees_to_use)
    return _SphericalHarmonics.apply(degrees_to_use, viewdirs.contiguous(), coeffs.contiguous())


class _SphericalHarmonics(Function):
    """autograd node over gsr_compute_sh_forward / gsr_compute_sh_backward (rasterizer/sh.py:62-97)."""

    @staticmethod
    def forward(ctx, degrees_to_use: int, viewdirs: Tensor, coeffs: Tensor):
        ctx.degrees_to_use = degrees_to_use
        ctx.degree = deg_from_sh(coeffs.shape[-2])
        ctx.save_for_backward(viewdirs)
        return _C.compute_sh_forward(coeffs.shape[0], ctx.degree, degrees_to_use, viewdirs, coeffs)

    @staticmethod
    def backward(ctx, v_colors: Tensor):
        (viewdirs,) = ctx.saved_tensors
        v_coeffs = _C.compute_sh_backward(v_colors.shape[0], ctx.degree, ctx.degrees_to_use, viewdirs,
                                          v_colors.contiguous())
        return None, None, v_coeffs
