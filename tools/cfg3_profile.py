"""A few cfg3-style training iterations (fused render + fused loss + one-launch Adam + statistics) of an object-centric
800x800 view with ~130 k Gaussians, for an ncu launch list (per-kernel share of one iteration)."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "gaussian-splatting-toolkit_b200"))
sys.path.insert(0, os.path.join(ROOT, "tests"))
from rasterizer.densify import DensifyStats  # noqa: E402
from rasterizer.fused import RenderAux, render_gaussians  # noqa: E402
from rasterizer.losses import l1_ssim_loss  # noqa: E402
from rasterizer.optim import GaussianOptimizers  # noqa: E402
from test_gpu_train_densify import H, W, _cameras, _gaussians  # noqa: E402

iters = int(sys.argv[1]) if len(sys.argv) > 1 else 6
cams, fx, fy = _cameras()
bg = torch.zeros(3, device="cuda")
teacher = {k: v.cuda() for k, v in _gaussians(200_000, 100, 0.004, 0.03, 0.55).items()}
p = {k: v.cuda().requires_grad_(True) for k, v in _gaussians(130_000, 101, 0.006, 0.03, 0.6).items()}


def render(q, cam, aux=None):
    rgb, _, _ = render_gaussians(q["means"], q["scales"], q["quats"], q["features_dc"], q["features_rest"], q["opacities"], cam[0], cam[1],
                                 fx, fy, W / 2.0, H / 2.0, H, W, 3, background=bg, render_depth=False, aux=aux)
    return torch.clamp(rgb, max=1.0)


with torch.no_grad():
    gts = [render(teacher, cams[i]) for i in range(8)]
opt = GaussianOptimizers(p)
stats, aux = DensifyStats(), RenderAux()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
for it in range(iters):
    if it == 2:
        torch.cuda.synchronize()
        e0.record()
    opt.zero_grad_all()
    loss = l1_ssim_loss(render(p, cams[it % 8], aux), gts[it % 8], 0.2)
    loss.backward()
    opt.optimizer_step_all()
    stats.update(aux.xys_grad, aux.radii, (H, W))
e1.record()
torch.cuda.synchronize()
print(f"ms/iter {e0.elapsed_time(e1) / max(1, iters - 2):.3f}  M={aux.num_intersects}")
