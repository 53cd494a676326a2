"""Generates tests/golden/ssim_kat.npz — vendor-free known-answer vectors for the SSIM of the photometric loss (SURVEY
8(f3); reference call site gs_toolkit/models/vanilla_gs.py:226,926-934: pytorch_msssim.SSIM(data_range=1.0,
size_average=True, channel=3)).  pytorch_msssim is absent from this image and from /root/reference, so the pin is an
INDEPENDENT float64 implementation of the published algorithm it implements (Wang, Bovik, Sheikh, Simoncelli, "Image
quality assessment: from error visibility to structural similarity", IEEE TIP 2004, eq. 13 with the paper's settings:
11x11 circular-symmetric Gaussian window, sigma 1.5, K1 = 0.01, K2 = 0.03, L = 1; window applied WITHOUT padding, mean
over the valid region and the channels, as pytorch_msssim does) written with numpy + scipy only: a full 2-D correlation
with the outer-product window (not the separable torch convolutions of oracle/ssim_ref.py), plus closed-form cases.

    python tests/golden/gen_golden_ssim.py
"""
import os

import numpy as np
from scipy.signal import correlate2d

HERE = os.path.dirname(os.path.abspath(__file__))
C1, C2 = 0.01 ** 2, 0.03 ** 2


def window():
    x = np.arange(11, dtype=np.float64) - 5
    g = np.exp(-x * x / (2 * 1.5 ** 2))
    g /= g.sum()
    return np.outer(g, g)


def ssim_wang(x, y):
    """x, y: [H, W, C] float64 in [0, 1] -> scalar mean SSIM."""
    w = window()
    vals = []
    for c in range(x.shape[2]):
        a, b = x[..., c], y[..., c]
        mu_a, mu_b = correlate2d(a, w, mode="valid"), correlate2d(b, w, mode="valid")
        s_aa = correlate2d(a * a, w, mode="valid") - mu_a * mu_a
        s_bb = correlate2d(b * b, w, mode="valid") - mu_b * mu_b
        s_ab = correlate2d(a * b, w, mode="valid") - mu_a * mu_b
        m = ((2 * mu_a * mu_b + C1) * (2 * s_ab + C2)) / ((mu_a ** 2 + mu_b ** 2 + C1) * (s_aa + s_bb + C2))
        vals.append(m.mean())
    return float(np.mean(vals))


def main():
    rng = np.random.default_rng(2004)
    out = {}
    cases = {
        "noise_32x40": (rng.random((32, 40, 3)), rng.random((32, 40, 3))),
        "noisy_copy_48x36": None,
        "blurred_edge_40x40": None,
        "dark_vs_bright_24x24": (np.full((24, 24, 3), 0.2), np.full((24, 24, 3), 0.7)),
    }
    base = rng.random((48, 36, 3))
    cases["noisy_copy_48x36"] = (base, np.clip(base + 0.05 * rng.standard_normal(base.shape), 0, 1))
    yy, xx = np.mgrid[0:40, 0:40]
    edge = (xx > 20).astype(np.float64)[..., None].repeat(3, axis=2)
    ramp = np.clip((xx - 15) / 10.0, 0, 1)[..., None].repeat(3, axis=2)
    cases["blurred_edge_40x40"] = (edge, ramp)
    for name, (a, b) in cases.items():
        a32, b32 = a.astype(np.float32), b.astype(np.float32)  # what the implementations under test receive
        out[name + "_x"], out[name + "_y"] = a32, b32
        out[name + "_ssim"] = np.float64(ssim_wang(a32.astype(np.float64), b32.astype(np.float64)))
        out[name + "_l1"] = np.float64(np.abs(a32.astype(np.float64) - b32.astype(np.float64)).mean())
    # closed form: two constant images a, b -> (2ab + C1) / (a^2 + b^2 + C1)  (variances vanish, the second factor is 1)
    assert abs(out["dark_vs_bright_24x24_ssim"] - (2 * 0.2 * 0.7 + C1) / (0.2 ** 2 + 0.7 ** 2 + C1)) < 1e-6
    np.savez_compressed(os.path.join(HERE, "ssim_kat.npz"), **out)
    for k, v in out.items():
        if k.endswith("_ssim"):
            print(f"{k:32s} {float(v):.12f}")


if __name__ == "__main__":
    main()
