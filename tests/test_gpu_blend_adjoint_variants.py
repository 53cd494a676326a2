"""The adjoint kernels of the 3-channel compositing — two-phase transposing (blend_bwd_tr.cu, default), warp prefix-scan
(blend_bwd_scan.cu, GSR_BWD_KERNEL=scan) and pixel-parallel (blend_bwd.cu, GSR_BWD_KERNEL=pixel) — must give the same
gradients up to the order of the FP32 sums.
The switch is read once per process, so each variant runs in its own interpreter."""
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
                 ("b", make_scene(8_000, 333, 222, 0.02, 0.4, margin=1.0, seed=32)),
                 ("c", make_scene(300, 40, 24, 0.05, 0.2, seed=1))):
    r = run_view_bindings(C, scene_to_torch(sc, "cuda"), sort_impl="gsr", binning="fast")
    for k in ("out_img", "final_Ts", "final_idx", "v_xy", "v_conic", "v_colors", "v_opacity"):
        out[name + "_" + k] = r[k].detach().cpu().numpy()
np.savez({path!r}, **out)
"""


def _run(bwd, tmp_path, **extra_env):
    path = str(tmp_path / ("_".join([bwd] + [f"{k}{v}" for k, v in extra_env.items()]) + ".npz"))
    env = dict(os.environ, GSR_BWD_KERNEL=bwd, **extra_env)
    subprocess.run([sys.executable, "-c", _SCRIPT.format(root=ROOT, path=path)], check=True, env=env)
    return np.load(path)


@pytest.mark.parametrize("variant", ["tr", "tr8", "tr32", "scan"])
def test_adjoint_variants_match_pixel_parallel(tmp_path, variant):
    """tr / tr8 / tr32 = two-phase transposing adjoint (blend_bwd_tr.cu; tr = 16 rows per group is the default kernel),
    scan = warp prefix-scan adjoint (blend_bwd_scan.cu)."""
    ref = _run("pixel", tmp_path)
    got = _run(variant, tmp_path)
    for name in ("a", "b", "c"):
        for k in ("out_img", "final_Ts", "final_idx"):
            assert np.array_equal(got[f"{name}_{k}"], ref[f"{name}_{k}"]), (name, k)
        for k in ("v_xy", "v_conic", "v_colors", "v_opacity"):
            a, b = got[f"{name}_{k}"].astype(np.float64), ref[f"{name}_{k}"].astype(np.float64)
            err = np.linalg.norm(a - b) / np.linalg.norm(b)
            print(f"[adjoint variants] {variant} scene {name} {k}: normwise rel {err:.2e}")
            assert err < 5e-6, (variant, name, k, err)


def test_block_masks_match_per_warp_tests(tmp_path):
    """The 16x16 kernels evaluate the warp-block reach of a staged record once, in the staging thread (block_mask_16,
    default), instead of once per warp (GSR_BLOCK_MASK=0, compact_survivors).  Both are conservative supersets of the
    pairs that pass the per-pixel alpha test, so the forward outputs must be BITWISE identical and the gradients equal
    up to the order of the FP32 atomics.  Scene "b" has Gaussians larger than a tile, "c" one partial tile row."""
    ref = _run("tr", tmp_path, GSR_BLOCK_MASK="0")
    got = _run("tr", tmp_path, GSR_BLOCK_MASK="1")
    for name in ("a", "b", "c"):
        for k in ("out_img", "final_Ts", "final_idx"):
            assert np.array_equal(got[f"{name}_{k}"], ref[f"{name}_{k}"]), (name, k)
        for k in ("v_xy", "v_conic", "v_colors", "v_opacity"):
            a, b = got[f"{name}_{k}"].astype(np.float64), ref[f"{name}_{k}"].astype(np.float64)
            err = np.linalg.norm(a - b) / np.linalg.norm(b)
            print(f"[block masks] scene {name} {k}: normwise rel {err:.2e}")
            assert err < 5e-6, (name, k, err)


_SH_SCRIPT = r"""
import sys, os, numpy as np, torch
sys.path.insert(0, {root!r}); sys.path.insert(0, os.path.join({root!r}, "gaussian-splatting-toolkit_b200"))
from rasterizer import cuda as C
g = torch.Generator().manual_seed(3)
out = {{}}
for n in (1, 127, 128, 100_003):
    dirs = torch.randn(n, 3, generator=g).cuda()
    coeffs = torch.randn(n, 16, 3, generator=g).cuda()
    for use in (0, 2, 3):
        out[f"{{n}}_{{use}}"] = C.compute_sh_forward(n, 3, use, dirs, coeffs).cpu().numpy()
np.savez({path!r}, **out)
"""


def test_sh_forward_tma_variant_is_bit_identical(tmp_path):
    """GSR_SH_TMA=1: the SH forward staged with TMA bulk copies (cp.async.bulk + mbarrier) gives bit-identical colours to the
    default cp.async variant — same arithmetic, only the staging differs; row counts that are not a multiple of the block."""
    res = {}
    for flag in ("0", "1"):
        path = str(tmp_path / f"sh_{flag}.npz")
        env = dict(os.environ, GSR_SH_TMA=flag)
        subprocess.run([sys.executable, "-c", _SH_SCRIPT.format(root=ROOT, path=path)], check=True, env=env)
        res[flag] = np.load(path)
    for k in res["0"].files:
        assert np.array_equal(res["0"][k], res["1"][k]), k
