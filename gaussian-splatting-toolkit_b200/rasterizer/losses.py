"""Fused photometric loss of the reference model:

    loss = (1 - ssim_lambda) * |gt - pred|.mean() + ssim_lambda * (1 - SSIM(gt, pred))

as in gs_toolkit/models/vanilla_gs.py:926-934 with pytorch_msssim.SSIM(data_range=1.0, size_average=True, channel=3)
(vanilla_gs.py:226), computed by one stencil kernel forward and one backward directly on the [H,W,3] images the
rasterizer produces (no permute / unsqueeze copies, no grouped convolutions)."""
from __future__ import annotations

import torch
from torch import Tensor
from torch.autograd import Function

from . import _lib
from .cuda import _Guard, _check_input, _ptr


def l1_ssim_loss(pred: Tensor, gt: Tensor, ssim_lambda: float = 0.2, return_terms: bool = False):
    """pred, gt: [H,W,3] float32 CUDA tensors, H, W >= 11.  Differentiable w.r.t. `pred`.
    With `return_terms` also returns (L1, SSIM) as detached scalars (for logging, vanilla_gs.py:883-898)."""
    if pred.shape != gt.shape or pred.dim() != 3 or pred.shape[-1] != 3:
        raise ValueError(f"pred and gt must both be [H,W,3], got {tuple(pred.shape)} and {tuple(gt.shape)}")
    loss, l1, ssim = _L1SSIM.apply(pred.contiguous(), gt.contiguous(), float(ssim_lambda))
    return (loss, l1, ssim) if return_terms else loss


class _L1SSIM(Function):
    @staticmethod
    def forward(ctx, pred, gt, ssim_lambda):
        _check_input(pred, "pred", torch.float32)
        _check_input(gt, "gt", torch.float32)
        H, W = pred.shape[0], pred.shape[1]
        if H < 11 or W < 11:
            raise ValueError("l1_ssim_loss needs images of at least 11 x 11 pixels (valid 11-tap window)")
        lib = _lib.load()
        n = lib.gsr_l1_ssim_num_partials(H, W)
        maps = torch.empty((3, H - 10, W - 10, 3), dtype=torch.float32, device=pred.device)
        partials = torch.empty((n, 2), dtype=torch.float32, device=pred.device)
        out = torch.empty((3,), dtype=torch.float32, device=pred.device)
        with _Guard(pred) as st:
            _lib.check(lib.gsr_l1_ssim_forward(H, W, ssim_lambda, _ptr(pred), _ptr(gt), _ptr(maps), _ptr(partials), _ptr(out),
                                               st), "l1_ssim_forward")
        loss, l1, ssim = out[0], out[1], out[2]
        ctx.ssim_lambda = ssim_lambda
        ctx.save_for_backward(pred, gt, maps)
        ctx.mark_non_differentiable(l1, ssim)
        return loss, l1, ssim

    @staticmethod
    def backward(ctx, v_loss, _v_l1, _v_ssim):
        pred, gt, maps = ctx.saved_tensors
        H, W = pred.shape[0], pred.shape[1]
        v_pred = torch.empty_like(pred)
        v_loss = v_loss.contiguous().float()
        with _Guard(pred) as st:
            _lib.check(_lib.load().gsr_l1_ssim_backward(H, W, ctx.ssim_lambda, _ptr(pred), _ptr(gt), _ptr(maps),
                                                        _ptr(v_loss), _ptr(v_pred), st), "l1_ssim_backward")
        return v_pred, None, None
