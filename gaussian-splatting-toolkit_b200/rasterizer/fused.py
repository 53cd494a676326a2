"""`render_gaussians` — the FUSED render operator.

One autograd node that takes the model's RAW parameters and a camera and returns (rgb, depth, alpha), doing inside
two per-Gaussian kernels and two per-tile kernels everything the reference model's `get_outputs` issues as ~25
separate torch ops and 5 operator calls per view (gs_toolkit/models/vanilla_gs.py:759-855):

    exp(scales), quats / |quats|, cat(features_dc, features_rest), viewdirs, spherical_harmonics, clamp(rgb + 0.5),
    sigmoid(opacities), project_gaussians, rasterize_gaussians (colour + alpha), rasterize_gaussians (depth)

Not part of the reference package's surface: a model has to opt in.  The numerics per pixel /
per Gaussian are those of the separate operators (same device functions); both `rasterize_mode`s of the model are
covered ("antialiased": opacity x EWA compensation, vanilla_gs.py:813-816).
"""
from __future__ import annotations

from typing import Optional, Tuple

import torch
from torch import Tensor
from torch.autograd import Function

from . import binning as _binning
from . import cuda as _C


class RenderAux:
    """Per-view by-products a model needs for densification (vanilla_gs.py:344-372): `radii` (int32) and `xys`
    (screen-space means; `xys.grad`-style statistics come from `xys_grad`, filled by the backward pass)."""

    def __init__(self):
        self.radii: Optional[Tensor] = None
        self.xys: Optional[Tensor] = None
        self.depths: Optional[Tensor] = None
        self.xys_grad: Optional[Tensor] = None
        self.num_intersects: int = 0


def render_gaussians(means3d: Tensor, scales: Tensor, quats: Tensor, features_dc: Tensor, features_rest: Tensor,
                     opacities: Tensor, viewmat: Tensor, projmat: Tensor, fx: float, fy: float, cx: float, cy: float,
                     img_height: int, img_width: int, degrees_to_use: int, background: Optional[Tensor] = None,
                     block_width: int = 16, render_depth: bool = True, glob_scale: float = 1.0,
                     clip_thresh: float = 0.01, aux: Optional[RenderAux] = None, rasterize_mode: str = "classic"
                     ) -> Tuple[Tensor, Optional[Tensor], Tensor]:
    """Render one view from RAW parameters: `scales` are log-scales, `quats` unnormalised (w,x,y,z), `opacities`
    logits [N,1], `features_dc` [N,3], `features_rest` [N,K-1,3].

    Returns (rgb [H,W,3], depth [H,W,1] or None, alpha [H,W,1]); rgb = sum_k w_k c_k + T * background (NOT clamped to
    1), depth = sum_k w_k z_k (NOT divided by alpha), alpha = 1 - T — i.e. exactly what the two rasterize_gaussians
    calls of the reference model return before its own epilogue (vanilla_gs.py:835-855).
    Differentiable w.r.t. means3d, scales, quats, features_dc, features_rest, opacities."""
    assert block_width > 1 and block_width <= 16, "block_width must be between 2 and 16"
    if rasterize_mode not in ("classic", "antialiased"):
        raise ValueError("Unknown rasterize_mode: %s" % rasterize_mode)   # vanilla_gs.py:817-818
    if features_rest.dim() != 3 or features_rest.shape[1] == 0:
        # sh_degree = 0 models colour with sigmoid(features_dc) (vanilla_gs.py:808), not with the SH band + 0.5 clamp this
        # operator fuses: refuse instead of silently rendering different colours than the separate operators
        raise ValueError("render_gaussians needs at least one SH band beyond dc (features_rest [N,K-1,3], K >= 4); "
                         "for sh_degree = 0 models use project_gaussians + rasterize_gaussians with sigmoid colours")
    if background is None:
        background = torch.ones(3, dtype=torch.float32, device=means3d.device)
    assert background.shape[0] == 3, "incorrect shape of background color tensor, expected shape 3"
    rgb, depth, alpha = _RenderGaussians.apply(
        means3d.contiguous(), scales.contiguous(), quats.contiguous(), features_dc.contiguous(),
        features_rest.contiguous(), opacities.contiguous(), viewmat.contiguous(), projmat.contiguous(), fx, fy, cx, cy,
        img_height, img_width, degrees_to_use, background.contiguous(), block_width, render_depth, glob_scale,
        clip_thresh, aux, rasterize_mode == "antialiased")
    return rgb, (depth[..., None] if render_depth else None), alpha[..., None]


class _RenderGaussians(Function):
    @staticmethod
    def forward(ctx, means3d, scales, quats, features_dc, features_rest, opacities, viewmat, projmat, fx, fy, cx, cy,
                img_height, img_width, degrees_to_use, background, block_width, render_depth, glob_scale, clip_thresh,
                aux, antialiased=False):
        n = means3d.shape[0]
        if n < 1 or means3d.shape[-1] != 3:
            raise ValueError(f"Invalid shape for means3d: {means3d.shape}")
        opac_raw = opacities.reshape(-1)
        rec, xys, depths, radii, conics, opac, mask, comp = _C.fused_preprocess_forward(
            means3d, scales, quats, opac_raw, features_dc.reshape(n, 3), features_rest, viewmat, projmat, glob_scale, fx, fy,
            cx, cy, img_height, img_width, block_width, degrees_to_use, clip_thresh, antialiased)
        # sync mode: one host read of M; async mode (rasterizer.binning): m is None, nothing synchronises
        m, ids_sorted, tile_bins = _binning.bin_gaussians(xys, depths, radii, conics, opac, img_height, img_width,
                                                          block_width)
        dev = means3d.device
        if m is None:
            m = -1  # unknown on the host (asynchronous binning): treated as non-empty, empty tiles blend to background
        if m == 0:
            rgb = torch.ones(img_height, img_width, 3, device=dev) * background
            depth = torch.zeros(img_height, img_width, device=dev)
            final_Ts = torch.ones(img_height, img_width, device=dev)
            final_idx = None
        else:
            rgb, depth, final_Ts, final_idx = _C.blend_packed_forward(img_height, img_width, block_width, ids_sorted,
                                                                      tile_bins, rec, background, render_depth)
            if depth is None:
                depth = torch.zeros(0, device=dev)
        if aux is not None:
            aux.radii, aux.xys, aux.depths, aux.num_intersects, aux.xys_grad = radii, xys, depths, m, None
        ctx.meta = (fx, fy, img_height, img_width, block_width, degrees_to_use, render_depth, glob_scale, m,
                    features_rest.shape[1] if features_rest.dim() == 3 else 0, tuple(features_dc.shape))
        ctx.aux = aux
        if m == 0:
            ctx.save_for_backward(means3d, scales, quats, opacities, viewmat, projmat)
        else:
            ctx.save_for_backward(means3d, scales, quats, opacities, viewmat, projmat, rec, radii, conics, mask,
                                  ids_sorted, tile_bins, background, final_Ts, final_idx,
                                  comp if comp is not None else torch.empty(0, device=dev))
        return rgb, depth, 1 - final_Ts

    @staticmethod
    def backward(ctx, v_rgb, v_depth, v_alpha):
        fx, fy, H, W, bw, degrees_to_use, render_depth, glob_scale, m, k_rest, dc_shape = ctx.meta
        saved = ctx.saved_tensors
        means3d, scales, quats, opacities, viewmat, projmat = saved[:6]
        n = means3d.shape[0]
        if m < 0:
            _binning.poll()  # non-blocking: raises if an earlier asynchronous call overflowed its pair buffers
        if m == 0:
            z = torch.zeros_like
            grads = (z(means3d), z(scales), z(quats), torch.zeros(n, 3, device=means3d.device),
                     torch.zeros(n, k_rest, 3, device=means3d.device), z(opacities))
            if ctx.aux is not None:
                ctx.aux.xys_grad = torch.zeros(n, 2, device=means3d.device)
        else:
            rec, radii, conics, mask, ids_sorted, tile_bins, background, final_Ts, final_idx, comp = saved[6:]
            v_rgb = v_rgb.contiguous()
            v_alpha = torch.zeros(H, W, device=v_rgb.device) if v_alpha is None else v_alpha.contiguous()
            v_d = v_depth.contiguous() if (render_depth and v_depth is not None) else None
            grad_rec = _C.blend_packed_backward(H, W, bw, ids_sorted, tile_bins, rec, background, final_Ts, final_idx,
                                                v_rgb, v_d, v_alpha)
            if ctx.aux is not None:
                ctx.aux.xys_grad = grad_rec[:, 0:2]
            v_means, v_scales, v_quats, v_opac, v_dc, v_rest = _C.fused_preprocess_backward(
                means3d, scales, quats, opacities.reshape(-1), k_rest, degrees_to_use, viewmat, projmat, glob_scale, fx, fy,
                H, W, radii, conics, mask, grad_rec, compensation=comp if comp.numel() else None)
            grads = (v_means, v_scales, v_quats, v_dc, v_rest, v_opac.view_as(opacities))
        v_means, v_scales, v_quats, v_dc, v_rest, v_opac = grads
        return (v_means, v_scales, v_quats, v_dc.reshape(dc_shape), v_rest, v_opac) + (None,) * 16
