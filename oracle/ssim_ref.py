"""TEST INFRASTRUCTURE — torch restatement of the photometric loss of the reference model
(gs_toolkit/models/vanilla_gs.py:926-934) including the third-party SSIM it calls:
pytorch_msssim.SSIM(data_range=1.0, size_average=True, channel=3) (vanilla_gs.py:226).  pytorch_msssim is NOT vendored
in /root/reference and is not installed in this image (pyproject.toml lists it unpinned); its published algorithm
(`ssim()` / `_ssim()` / `gaussian_filter()` / `_fspecial_gauss_1d()` of pytorch_msssim 1.0.0) is restated here:
separable 11-tap Gaussian window (sigma 1.5), VALID padding, K = (0.01, 0.03), mean over channels and pixels.
Gradients come from torch autograd through this restatement.  Parity unpinned against the package itself (absent);
anchored on the reference's call site and defaults."""
import torch
import torch.nn.functional as F


def gaussian_window(size=11, sigma=1.5, device="cpu", dtype=torch.float32):
    coords = torch.arange(size, dtype=dtype, device=device) - size // 2
    g = torch.exp(-(coords ** 2) / (2 * sigma ** 2))
    return g / g.sum()


def ssim(x, y, data_range=1.0, win_size=11, sigma=1.5, k=(0.01, 0.03)):
    """x, y: [B,C,H,W].  Returns the scalar pytorch_msssim.ssim(x, y, data_range, size_average=True)."""
    c = x.shape[1]
    g = gaussian_window(win_size, sigma, x.device, x.dtype)
    wv, wh = g.view(1, 1, -1, 1).repeat(c, 1, 1, 1), g.view(1, 1, 1, -1).repeat(c, 1, 1, 1)

    def filt(t):
        return F.conv2d(F.conv2d(t, wv, groups=c), wh, groups=c)

    c1, c2 = (k[0] * data_range) ** 2, (k[1] * data_range) ** 2
    mu1, mu2 = filt(x), filt(y)
    s1, s2, s12 = filt(x * x) - mu1 * mu1, filt(y * y) - mu2 * mu2, filt(x * y) - mu1 * mu2
    cs = (2 * s12 + c2) / (s1 + s2 + c2)
    ssim_map = ((2 * mu1 * mu2 + c1) / (mu1 * mu1 + mu2 * mu2 + c1)) * cs
    return torch.flatten(ssim_map, 2).mean(-1).mean()


def l1_ssim_loss(pred_hwc, gt_hwc, ssim_lambda=0.2):
    """The reference's loss on [H,W,3] images (vanilla_gs.py:926-934; ssim(gt, pred) argument order as there)."""
    l1 = torch.abs(gt_hwc - pred_hwc).mean()
    s = ssim(gt_hwc.permute(2, 0, 1)[None], pred_hwc.permute(2, 0, 1)[None])
    return (1 - ssim_lambda) * l1 + ssim_lambda * (1 - s), l1, s
