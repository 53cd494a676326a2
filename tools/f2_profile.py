"""Workload for the ncu pass over the f2 kernels (tools/gpu_f2.sh): one Adam step over 59 floats x 1 M Gaussians and one
densification refinement of 1 M Gaussians, a few times each."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "gaussian-splatting-toolkit_b200"))
from rasterizer.densify import DensifyConfig, DensifyStats, refinement_after  # noqa: E402
from rasterizer.optim import DEFAULT_LRS, GaussianOptimizers  # noqa: E402

n = 1_000_000
g = torch.Generator(device="cuda").manual_seed(0)
R = lambda *s: torch.rand(*s, device="cuda", generator=g)
G = lambda *s: torch.randn(*s, device="cuda", generator=g)
base = {"means": G(n, 3) * 2, "scales": (R(n, 1) * (np.log(0.8) - np.log(0.001)) + np.log(0.001)) + torch.log(0.5 + 0.5 * R(n, 3)),
        "quats": G(n, 4), "features_dc": R(n, 3), "features_rest": G(n, 15, 3) * 0.05, "opacities": G(n, 1) * 2}
for rep in range(3):
    params = {k: v.clone().requires_grad_(True) for k, v in base.items()}
    opt = GaussianOptimizers(params, DEFAULT_LRS)
    for _ in range(2):
        opt.optimizer_step_all({k: torch.randn_like(v) * 1e-3 for k, v in params.items()})
    stats = DensifyStats()
    for _ in range(2):
        stats.update(G(n, 2) * 4e-6, torch.randint(-2, 200, (n,), device="cuda", dtype=torch.int32, generator=g), (1080, 1920))
    info = refinement_after(params, opt, stats, DensifyConfig(), 3500, 200, (1080, 1920))
    torch.cuda.synchronize()
print(info)
