"""Host-side cost per operator call (tiny scene => GPU time negligible): this package vs the reference extension."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "gaussian-splatting-toolkit_b200"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import torch
import rasterizer
from rasterizer import cuda as C
from rasterizer.sh import spherical_harmonics
from rasterizer.synthetic import make_scene, scene_to_torch
from oracle.build_ref import load_ref
from pipelines import run_view_bindings, run_view_public

s = scene_to_torch(make_scene(2000, 96, 64, 0.03, 0.2, seed=1), "cuda")
ref = load_ref()

def wall(fn, n=300, warm=30):
    for _ in range(warm): fn()
    torch.cuda.synchronize(); t = time.perf_counter()
    for _ in range(n): fn()
    torch.cuda.synchronize(); return (time.perf_counter() - t) / n * 1e6

N = 2000
viewdirs = (s["means3d"] - s["cam_pos"][None]).contiguous()
print("compute_sh_forward   ours %.1f us" % wall(lambda: C.compute_sh_forward(N, 3, 3, viewdirs, s["sh_coeffs"])), end="")
if ref: print("   ref %.1f us" % wall(lambda: ref.compute_sh_forward(N, 3, 3, viewdirs, s["sh_coeffs"])))
args = (N, s["means3d"], s["scales"], 1.0, s["quats"], s["viewmat"], s["projmat"], s["fx"], s["fy"], s["cx"], s["cy"], 64, 96, 16, 0.01)
print("project_forward      ours %.1f us" % wall(lambda: C.project_gaussians_forward(*args)), end="")
if ref: print("   ref %.1f us" % wall(lambda: ref.project_gaussians_forward(*args)))
print("view fwd+bwd (bindings, ref binning via torch ops) ours %.1f us" % wall(lambda: run_view_bindings(C, s, sort_impl="torch"), 100, 10), end="")
if ref: print("   ref %.1f us" % wall(lambda: run_view_bindings(ref, s, sort_impl="torch"), 100, 10))
print("view fwd+bwd (bindings, fast binning)              ours %.1f us" % wall(lambda: run_view_bindings(C, s, sort_impl="gsr", binning="fast"), 100, 10))
print("view fwd+bwd (public autograd API)                 ours %.1f us" % wall(lambda: run_view_public(s), 100, 10))
cov3d, xys, depths, radii, conics, comp, nth = C.project_gaussians_forward(*args)
opac = s["opacities"].contiguous()
print("bin_gaussians_fast   ours %.1f us" % wall(lambda: C.bin_gaussians_fast(xys, depths, radii, conics, opac, 64, 96, 16)))
