"""EWA projection of 3D Gaussians — same surface as the reference `rasterizer.project_gaussians`
(rasterizer/project_gaussians.py:12-232)."""
from __future__ import annotations

from typing import Tuple

import torch
from torch import Tensor
from torch.autograd import Function

from . import cuda as _C


def project_gaussians(means3d: Tensor, scales: Tensor, glob_scale: float, quats: Tensor, viewmat: Tensor,
                      projmat: Tensor, fx: float, fy: float, cx: float, cy: float, img_height: int, img_width: int,
                      block_width: int, clip_thresh: float = 0.01
                      ) -> Tuple[Tensor, Tensor, Tensor, Tensor, Tensor, Tensor, Tensor]:
    """Project N Gaussians (means3d [N,3], scales [N,3], quats [N,4] wxyz) with `viewmat` (world->camera) and
    `projmat` (= P @ viewmat).  Differentiable w.r.t. means3d, scales, quats.

    Returns (xys [N,2], depths [N], radii [N] i32, conics [N,3], compensation [N], num_tiles_hit [N] i32,
    cov3d [N,6]) — the order of rasterizer/project_gaussians.py:153."""
    assert block_width > 1 and block_width <= 16, "block_width must be between 2 and 16"
    return _ProjectGaussians.apply(means3d.contiguous(), scales.contiguous(), glob_scale, quats.contiguous(),
                                   viewmat.contiguous(), projmat.contiguous(), fx, fy, cx, cy, img_height,
                                   img_width, block_width, clip_thresh)


class _ProjectGaussians(Function):
    """autograd node over gsr_project_gaussians_forward / _backward."""

    @staticmethod
    def forward(ctx, means3d, scales, glob_scale, quats, viewmat, projmat, fx, fy, cx, cy, img_height, img_width,
                block_width, clip_thresh=0.01):
        num_points = means3d.shape[-2]
        if num_points < 1 or means3d.shape[-1] != 3:
            raise ValueError(f"Invalid shape for means3d: {means3d.shape}")
        cov3d, xys, depths, radii, conics, compensation, num_tiles_hit = _C.project_gaussians_forward(
            num_points, means3d, scales, glob_scale, quats, viewmat, projmat, fx, fy, cx, cy, img_height,
            img_width, block_width, clip_thresh)
        ctx.meta = (num_points, glob_scale, fx, fy, cx, cy, img_height, img_width)
        ctx.save_for_backward(means3d, scales, quats, viewmat, projmat, cov3d, radii, conics, compensation)
        ctx.mark_non_differentiable(radii, num_tiles_hit)
        return xys, depths, radii, conics, compensation, num_tiles_hit, cov3d

    @staticmethod
    def backward(ctx, v_xys, v_depths, v_radii, v_conics, v_compensation, v_num_tiles_hit, v_cov3d):
        means3d, scales, quats, viewmat, projmat, cov3d, radii, conics, compensation = ctx.saved_tensors
        num_points, glob_scale, fx, fy, cx, cy, img_height, img_width = ctx.meta

        def dense(g, like):
            if g is None:
                return torch.zeros(like.shape, dtype=torch.float32, device=like.device)
            return g.contiguous()

        # v_cov3d (a gradient flowing into the returned cov3d) is ignored, as in the reference
        # (project_gaussians.py:155-203 never forwards it).
        _, _, v_mean3d, v_scale, v_quat = _C.project_gaussians_backward(
            num_points, means3d, scales, glob_scale, quats, viewmat, projmat, fx, fy, cx, cy, img_height,
            img_width, cov3d, radii, conics, compensation, dense(v_xys, means3d[:, :2]),
            dense(v_depths, compensation), dense(v_conics, conics), dense(v_compensation, compensation),
            need_cov_grads=False)
        return (v_mean3d, v_scale, None, v_quat, None, None, None, None, None, None, None, None, None, None)
