"""cProfile of the public autograd API path on a tiny scene (GPU time negligible): where the HOST time of a view goes."""
import cProfile, os, pstats, sys, io
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "gaussian-splatting-toolkit_b200"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import torch
import rasterizer
from rasterizer import binning
from rasterizer.synthetic import make_scene, scene_to_torch
from pipelines import run_view_public

mode = sys.argv[1] if len(sys.argv) > 1 else "sync"
binning.set_binning_mode(mode)
s = scene_to_torch(make_scene(2000, 96, 64, 0.03, 0.2, seed=1), "cuda")
for _ in range(30):
    run_view_public(s)
torch.cuda.synchronize()
import time
t = time.perf_counter()
for _ in range(200):
    run_view_public(s)
torch.cuda.synchronize()
print(f"[{mode}] public API view fwd+bwd: {(time.perf_counter() - t) / 200 * 1e6:.1f} us host wall per view")
pr = cProfile.Profile()
pr.enable()
for _ in range(200):
    run_view_public(s)
torch.cuda.synchronize()
pr.disable()
st = io.StringIO()
pstats.Stats(pr, stream=st).sort_stats("tottime").print_stats(28)
print(st.getvalue()[:6000])
