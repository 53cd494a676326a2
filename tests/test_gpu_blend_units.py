"""EXPERIMENTAL (branch r02-subwarp-units): the blend kernels with sub-warp units
(GSR_BLEND_UNITS = 2 | 4) against the default kernels.  The switch is read once per process, so each variant runs
in its own interpreter and dumps its outputs; the forward image / T / final_idx must be BITWISE equal (same per-pixel
arithmetic in the same order), the gradients equal up to the order of the FP32 sums."""
import os
import subprocess
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

_SCRIPT = r"""
import sys, os, numpy as np, torch
sys.path.insert(0, {root!r}); sys.path.insert(0, os.path.join({root!r}, "gaussian-splatting-toolkit_b200")); sys.path.insert(0, os.path.join({root!r}, "tests"))
from pipelines import run_view_bindings
from rasterizer import cuda as C
from rasterizer.synthetic import make_scene, scene_to_torch
out = {{}}
for name, sc in (("a", make_scene(60_000, 500, 300, 0.004, 0.05, margin=1.1, seed=31)),
                 ("b", make_scene(8_000, 333, 222, 0.02, 0.4, margin=1.0, seed=32))):
    r = run_view_bindings(C, scene_to_torch(sc, "cuda"), sort_impl="gsr", binning="fast")
    for k in ("out_img", "final_Ts", "final_idx", "v_xy", "v_conic", "v_colors", "v_opacity"):
        out[name + "_" + k] = r[k].detach().cpu().numpy()
import time; t0 = time.time()
sc = scene_to_torch(make_scene(1_000_000, 1920, 1080, 0.002, 0.02, margin=1.1, seed=0), "cuda")
for _ in range(3): run_view_bindings(C, sc, sort_impl="gsr", binning="fast")
torch.cuda.synchronize(); e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10): run_view_bindings(C, sc, sort_impl="gsr", binning="fast")
e1.record(); torch.cuda.synchronize()
out["ms_per_view_cfg2"] = e0.elapsed_time(e1) / 10
np.savez({path!r}, **out)
"""


def _run(units, tmp_path, bwd="pixel"):
    path = str(tmp_path / f"units{units}_{bwd}.npz")
    env = dict(os.environ, GSR_BLEND_UNITS=str(units), GSR_BWD_KERNEL=bwd)
    subprocess.run([sys.executable, "-c", _SCRIPT.format(root=ROOT, path=path)], check=True, env=env)
    return np.load(path)


def test_gaussian_parallel_scan_adjoint_matches_pixel_parallel(tmp_path):
    """GSR_BWD_KERNEL=scan (blend_bwd_scan.cu: lane = Gaussian, warp prefix scans carry the per-pixel state) against the
    default pixel-parallel adjoint: same gradients up to the order of the FP32 sums."""
    ref = _run(1, tmp_path)
    got = _run(1, tmp_path, bwd="scan")
    for name in ("a", "b"):
        for k in ("out_img", "final_Ts", "final_idx"):
            assert np.array_equal(got[f"{name}_{k}"], ref[f"{name}_{k}"]), (name, k)
        for k in ("v_xy", "v_conic", "v_colors", "v_opacity"):
            a, b = got[f"{name}_{k}"].astype(np.float64), ref[f"{name}_{k}"].astype(np.float64)
            err = np.linalg.norm(a - b) / np.linalg.norm(b)
            print(f"[scan adjoint] scene {name} {k}: normwise rel vs pixel-parallel {err:.2e}")
            assert err < 5e-6, (name, k, err)
    print(f"[scan adjoint] cfg2 view (bindings harness) {float(got['ms_per_view_cfg2']):.3f} ms vs pixel-parallel "
          f"{float(ref['ms_per_view_cfg2']):.3f} ms")


def test_subwarp_units_match_default_kernels(tmp_path):
    ref = _run(1, tmp_path)
    for units in (2, 4):
        got = _run(units, tmp_path)
        for name in ("a", "b"):
            for k in ("out_img", "final_Ts", "final_idx"):
                assert np.array_equal(got[f"{name}_{k}"], ref[f"{name}_{k}"]), (units, name, k)
            for k in ("v_xy", "v_conic", "v_colors", "v_opacity"):
                a, b = got[f"{name}_{k}"].astype(np.float64), ref[f"{name}_{k}"].astype(np.float64)
                err = np.linalg.norm(a - b) / np.linalg.norm(b)
                assert err < 2e-6, (units, name, k, err)
        print(f"[units] GSR_BLEND_UNITS={units}: cfg2 view (bindings harness) {float(got['ms_per_view_cfg2']):.3f} ms "
              f"vs default {float(ref['ms_per_view_cfg2']):.3f} ms")
