"""`GaussianRasterizer` / `GaussianRasterizationSettings` — an API-shape façade in the style of the Inria
diff-gaussian-rasterization package, layered on the three operators of this package.

The reference toolkit itself never imports such classes (its models call the functional gsplat-0.1
surface `project_gaussians` / `spherical_harmonics` / `rasterize_gaussians`, gs_toolkit/models/vanilla_gs.py:765-855);
the façade exists because BASELINE.json's north star names this class pair, and for users coming from Inria-style
code.  The arithmetic is the toolkit's (gsplat conventions: pixel centres at integer coordinates, 0.3 px blur, the
reference's alpha thresholds), NOT bit-compatible with Inria's CUDA rasterizer.

Conventions accepted here are Inria's: `viewmatrix` and `projmatrix` are the TRANSPOSED (row-vector) world-to-view and
full-projection matrices; `rotations` are (w,x,y,z); `means2D` is a zero tensor whose `.grad` receives the
screen-space mean gradients (densification statistic).
"""
from __future__ import annotations

from typing import NamedTuple, Optional

import torch
from torch import Tensor, nn

from .project_gaussians import project_gaussians
from .rasterize import rasterize_gaussians
from .sh import spherical_harmonics


class GaussianRasterizationSettings(NamedTuple):
    image_height: int
    image_width: int
    tanfovx: float
    tanfovy: float
    bg: Tensor
    scale_modifier: float
    viewmatrix: Tensor
    projmatrix: Tensor
    sh_degree: int
    campos: Tensor
    prefiltered: bool = False
    debug: bool = False
    block_width: int = 16


class GaussianRasterizer(nn.Module):
    def __init__(self, raster_settings: GaussianRasterizationSettings):
        super().__init__()
        self.raster_settings = raster_settings

    def forward(self, means3D: Tensor, means2D: Optional[Tensor], opacities: Tensor, shs: Optional[Tensor] = None,
                colors_precomp: Optional[Tensor] = None, scales: Optional[Tensor] = None,
                rotations: Optional[Tensor] = None, cov3D_precomp: Optional[Tensor] = None,
                return_depth_alpha: bool = False):
        """Returns (color [3,H,W], radii [N]) — plus (depth [1,H,W], alpha [1,H,W]) when `return_depth_alpha`.
        Gradients flow to means3D, scales, rotations, shs / colors_precomp, opacities, and (as the screen-space
        mean gradient) to means2D[:, :2]."""
        rs = self.raster_settings
        if (shs is None) == (colors_precomp is None):
            raise Exception("Please provide excatly one of either SHs or precomputed colors!")
        if cov3D_precomp is not None or scales is None or rotations is None:
            raise NotImplementedError("precomputed 3D covariances are not supported: pass scales and rotations")
        H, W, bw = rs.image_height, rs.image_width, rs.block_width
        fx, fy = 0.5 * W / rs.tanfovx, 0.5 * H / rs.tanfovy
        viewmat = rs.viewmatrix.t().contiguous()
        projmat = rs.projmatrix.t().contiguous()
        xys, depths, radii, conics, comp, num_tiles_hit, cov3d = project_gaussians(
            means3D, scales, rs.scale_modifier, rotations, viewmat, projmat, fx, fy, 0.5 * W, 0.5 * H, H, W, bw)
        if means2D is not None:
            xys = xys + means2D[..., :2]  # means2D is zeros; its .grad receives d loss / d xys
        if shs is not None:
            viewdirs = means3D.detach() - rs.campos.reshape(1, 3)
            colors = torch.clamp(spherical_harmonics(rs.sh_degree, viewdirs, shs) + 0.5, min=0.0)
        else:
            colors = colors_precomp
        opac = opacities.reshape(-1, 1)
        img, alpha = rasterize_gaussians(xys, depths, radii, conics, num_tiles_hit, colors, opac, H, W, bw,
                                         background=rs.bg, return_alpha=True)
        color = img.permute(2, 0, 1)
        if not return_depth_alpha:
            return color, radii
        depth = rasterize_gaussians(xys, depths, radii, conics, num_tiles_hit, depths[:, None].repeat(1, 3), opac, H, W,
                                    bw, background=torch.zeros(3, device=img.device))[..., 0:1]
        return color, radii, depth.permute(2, 0, 1), alpha[None]
