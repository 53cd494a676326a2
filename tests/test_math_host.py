"""CPU parity of the product SOURCE of the per-Gaussian maths against the oracle.

csrc/project_math.cuh (EWA projection and its adjoint) and csrc/sh_math.cuh (SH basis) are __host__ __device__;
tests/host_cull/host_math.cu wraps them for the host (built here with nvcc, no GPU).  The same seeded scenes go through
that build and through the oracle (oracle/gsr_oracle.c, itself pinned to the reference's _torch_impl and to the compiled
reference extension): integer outputs must be identical, FP32 outputs agree to a few ulp (two IEEE builds of the same
expressions; contraction into FMAs may differ).  The device build of the same source is what `-m gpu` tests compare with
the oracle and with the live reference extension."""
import ctypes
import math
import os
import shutil
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "gaussian-splatting-toolkit_b200"))
SRC = os.path.join(ROOT, "tests", "host_cull", "host_math.cu")
CSRC = os.path.join(ROOT, "gaussian-splatting-toolkit_b200", "csrc")


@pytest.fixture(scope="module")
def lib(tmp_path_factory):
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(nvcc):
        pytest.skip("nvcc not available")
    out = str(tmp_path_factory.mktemp("host_math") / "libhost_math.so")
    subprocess.run([nvcc, "-gencode", "arch=compute_100a,code=sm_100a", "-O2", "-std=c++17", "--expt-relaxed-constexpr",
                    "-shared", "-Xcompiler", "-fPIC", "-I", CSRC, "-o", out, SRC], check=True)
    return ctypes.CDLL(out)


def _p(a):
    return a.ctypes.data_as(ctypes.c_void_p)


def _rotated_view():
    """a camera that is neither axis-aligned nor at the origin"""
    ax, ay = 0.21, -0.17
    Rx = np.array([[1, 0, 0], [0, math.cos(ax), -math.sin(ax)], [0, math.sin(ax), math.cos(ax)]])
    Ry = np.array([[math.cos(ay), 0, math.sin(ay)], [0, 1, 0], [-math.sin(ay), 0, math.cos(ay)]])
    V = np.eye(4)
    V[:3, :3] = Rx @ Ry
    V[:3, 3] = [0.3, -0.2, 0.5]
    return V.astype(np.float32)


def _rel(a, b):
    a, b = a.astype(np.float64), b.astype(np.float64)
    return np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-30)


@pytest.mark.parametrize("case", ["identity_bw16", "rotated_bw8"])
def test_projection_source_matches_oracle_on_cpu(lib, case):
    from oracle import oracle as orc
    from rasterizer.synthetic import make_scene

    if case == "identity_bw16":
        s = make_scene(20_000, 640, 360, 0.004, 0.08, margin=1.2, seed=4)
    else:
        s = make_scene(20_000, 333, 222, 0.01, 0.3, margin=1.3, seed=5, block_width=8, viewmat=_rotated_view())
    n, H, W, bw = s["means3d"].shape[0], s["img_height"], s["img_width"], s["block_width"]
    f = ctypes.c_float
    ref = orc.project_forward(s["means3d"], s["scales"], s["glob_scale"], s["quats"], s["viewmat"], s["projmat"], s["fx"],
                              s["fy"], s["cx"], s["cy"], H, W, bw, s["clip_thresh"])
    cov3d, xys, depths = np.zeros((n, 6), np.float32), np.zeros((n, 2), np.float32), np.zeros(n, np.float32)
    radii, conics, comp, tiles = np.zeros(n, np.int32), np.zeros((n, 3), np.float32), np.zeros(n, np.float32), np.zeros(n, np.int32)
    lib.host_project_forward(n, _p(s["means3d"]), _p(s["scales"]), f(s["glob_scale"]), _p(s["quats"]), _p(s["viewmat"]),
                             _p(s["projmat"]), f(s["fx"]), f(s["fy"]), f(s["cx"]), f(s["cy"]), H, W, bw, f(s["clip_thresh"]),
                             _p(cov3d), _p(xys), _p(depths), _p(radii), _p(conics), _p(comp), _p(tiles))
    r_cov3d, r_xys, r_depths, r_radii, r_conics, r_comp, r_tiles = ref
    vis = r_radii > 0
    assert vis.sum() > 0.5 * n and (~vis).sum() > 100  # both visible and culled Gaussians are present
    # integers: identical except where the radius' ceil / the box truncation sits within an ulp of an integer
    bad_r, bad_t = int((radii != r_radii).sum()), int((tiles != r_tiles).sum())
    print(f"[host projection {case}] radii mismatches {bad_r}, num_tiles_hit mismatches {bad_t} of {n}")
    assert bad_r <= 2 and bad_t <= 2
    same = vis & (radii == r_radii) & (tiles == r_tiles)
    for name, a, b in (("xys", xys, r_xys), ("depths", depths, r_depths), ("conics", conics, r_conics),
                       ("compensation", comp, r_comp), ("cov3d", cov3d, r_cov3d)):
        err = _rel(a[same], b[same])
        print(f"[host projection {case}] {name}: normwise rel {err:.2e}")
        assert err < 2e-6, (name, err)
    assert np.array_equal(radii[~vis & (radii == r_radii)], r_radii[~vis & (radii == r_radii)])

    # adjoint with every upstream gradient live (v_xy, v_depth, v_conic, v_compensation)
    g = np.random.default_rng(7)
    v_xy, v_depth = g.normal(size=(n, 2)).astype(np.float32), g.normal(size=n).astype(np.float32)
    v_conic, v_comp = g.normal(size=(n, 3)).astype(np.float32), g.normal(size=n).astype(np.float32)
    _, _, r_vm, r_vs, r_vq = orc.project_backward(s["means3d"], s["scales"], s["glob_scale"], s["quats"], s["viewmat"],
                                                  s["projmat"], s["fx"], s["fy"], s["cx"], s["cy"], H, W, r_cov3d, r_radii,
                                                  r_conics, r_comp, v_xy, v_depth, v_conic, v_comp)
    v_mean, v_scale, v_quat = np.zeros((n, 3), np.float32), np.zeros((n, 3), np.float32), np.zeros((n, 4), np.float32)
    lib.host_project_backward(n, _p(s["means3d"]), _p(s["scales"]), f(s["glob_scale"]), _p(s["quats"]), _p(s["viewmat"]),
                              _p(s["projmat"]), f(s["fx"]), f(s["fy"]), H, W, _p(r_cov3d), _p(r_radii), _p(r_conics),
                              _p(r_comp), _p(v_xy), _p(v_depth), _p(v_conic), _p(v_comp), _p(v_mean), _p(v_scale), _p(v_quat))
    for name, a, b in (("v_mean3d", v_mean, r_vm), ("v_scale", v_scale, r_vs), ("v_quat", v_quat, r_vq)):
        err = _rel(a, b)
        print(f"[host projection {case}] {name}: normwise rel {err:.2e}")
        assert err < 1e-5, (name, err)
        assert not a[~vis].any()  # culled Gaussians get zero gradients


def test_sh_basis_source_matches_oracle_on_cpu(lib):
    from oracle import oracle as orc

    g = np.random.default_rng(2)
    n = 5000
    dirs = g.normal(size=(n, 3)).astype(np.float32) * g.uniform(0.1, 20.0, (n, 1)).astype(np.float32)
    for degree in (0, 1, 2, 3, 4):
        K = (degree + 1) ** 2
        coeffs = g.normal(size=(n, K, 3)).astype(np.float32)
        for use in range(degree + 1):
            ref = orc.sh_forward(use, dirs, coeffs)
            got = np.zeros((n, 3), np.float32)
            lib.host_sh_forward(n, K, use, _p(dirs), _p(coeffs), _p(got))
            err = _rel(got, ref)
            assert err < 2e-6, (degree, use, err)
