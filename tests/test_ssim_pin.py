"""Pins the SSIM restatement (oracle/ssim_ref.py) — the checker of the fused L1 + SSIM kernels — to vendor-free
known-answer vectors: an independent numpy / scipy float64 implementation of Wang et al. 2004 (full 2-D window
correlation, tests/golden/gen_golden_ssim.py) and closed forms.  CPU only."""
import os

import numpy as np
import torch

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "ssim_kat.npz")
CASES = ("noise_32x40", "noisy_copy_48x36", "blurred_edge_40x40", "dark_vs_bright_24x24")


def test_ssim_restatement_matches_independent_wang2004_vectors():
    from oracle.ssim_ref import l1_ssim_loss, ssim

    z = np.load(GOLD)
    for name in CASES:
        x, y = torch.from_numpy(z[name + "_x"]).double(), torch.from_numpy(z[name + "_y"]).double()
        s = float(ssim(x.permute(2, 0, 1)[None], y.permute(2, 0, 1)[None]))
        assert abs(s - float(z[name + "_ssim"])) < 1e-10, (name, s, float(z[name + "_ssim"]))
        loss, l1, s2 = l1_ssim_loss(y, x, 0.2)  # (pred, gt): SSIM is symmetric in its two images
        assert abs(float(l1) - float(z[name + "_l1"])) < 1e-12
        assert abs(float(loss) - (0.8 * float(z[name + "_l1"]) + 0.2 * (1 - float(z[name + "_ssim"])))) < 1e-10


def test_ssim_closed_forms():
    from oracle.ssim_ref import ssim

    g = torch.Generator().manual_seed(0)
    x = torch.rand(1, 3, 30, 26, generator=g, dtype=torch.float64)
    assert abs(float(ssim(x, x)) - 1.0) < 1e-12  # identity
    a, b = 0.35, 0.6
    ca, cb = torch.full((1, 3, 20, 20), a, dtype=torch.float64), torch.full((1, 3, 20, 20), b, dtype=torch.float64)
    assert abs(float(ssim(ca, cb)) - (2 * a * b + 1e-4) / (a * a + b * b + 1e-4)) < 1e-12
    assert abs(float(ssim(x, 0.5 * x)) - float(ssim(0.5 * x, x))) < 1e-14  # symmetry
