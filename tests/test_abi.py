"""CPU-only checks of the drop-in boundary: the C-ABI library loads, exports every symbol declared in
include/gsr_b200.h, the Python package mirrors the reference's public surface, and argument errors are
raised before anything touches a GPU."""
import ctypes
import inspect
import os
import re

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "gsr_b200.h")


def _declared_symbols():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"GSR_API\s+[\w\s\*]+?\b(gsr_\w+)\s*\(", src)))


@pytest.fixture(scope="module")
def lib():
    from rasterizer import _lib

    if not os.path.exists(_lib.LIB_PATH):
        _lib.build_library()
    return _lib.load()


def test_header_declares_expected_entry_points():
    syms = _declared_symbols()
    for name in ("gsr_compute_sh_forward", "gsr_compute_sh_backward", "gsr_project_gaussians_forward",
                 "gsr_project_gaussians_backward", "gsr_compute_cov2d_bounds", "gsr_map_gaussian_to_intersects",
                 "gsr_get_tile_bin_edges", "gsr_rasterize_forward", "gsr_rasterize_backward",
                 "gsr_nd_rasterize_forward", "gsr_nd_rasterize_backward"):
        assert name in syms  # one per reference binding, csrc/ext.cpp:6-17


def test_library_exports_every_declared_symbol(lib):
    from rasterizer import _lib

    raw = ctypes.CDLL(_lib.LIB_PATH)
    declared = _declared_symbols()
    assert len(declared) >= 18
    for name in declared:
        assert hasattr(raw, name), f"{name} declared in gsr_b200.h but not exported"
        assert name in _lib.SIGNATURES, f"{name} has no ctypes prototype in rasterizer/_lib.py"
    assert set(_lib.SIGNATURES) == set(declared)


def test_library_identification(lib):
    assert lib.gsr_version().decode().startswith("0.1.2")
    assert lib.gsr_built_for_sm() == 100
    assert lib.gsr_cumsum_workspace_bytes(1000) > 0
    assert lib.gsr_sort_workspace_bytes(1000) > 0


def test_argument_errors_do_not_need_a_gpu(lib):
    # block_width outside [2,16] -> GSR_ERR_INVALID_ARGUMENT with a message (no launch attempted)
    rc = lib.gsr_rasterize_forward(16, 16, 17, 1, None, None, None, None, None, None, None, None, None, None, None)
    assert rc == -1 and b"block_width" in lib.gsr_last_error()
    rc = lib.gsr_compute_sh_forward(10, 5, 0, None, None, None, None)
    assert rc == -4
    rc = lib.gsr_compute_sh_forward(0, 3, 3, None, None, None, None)
    assert rc == 0  # empty input is a no-op


def test_python_surface_matches_reference():
    import rasterizer
    from rasterizer import cuda as C
    import importlib

    pg_mod = importlib.import_module("rasterizer.project_gaussians")
    rz_mod = importlib.import_module("rasterizer.rasterize")
    sh_mod = importlib.import_module("rasterizer.sh")
    from rasterizer._torch_impl import quat_to_rotmat

    assert rasterizer.__version__ == "0.1.2"
    for name in ("project_gaussians", "rasterize_gaussians", "spherical_harmonics", "bin_and_sort_gaussians",
                 "compute_cumulative_intersects", "compute_cov2d_bounds", "get_tile_bin_edges",
                 "map_gaussian_to_intersects", "ProjectGaussians", "RasterizeGaussians", "BinAndSortGaussians",
                 "ComputeCumulativeIntersects", "ComputeCov2dBounds", "GetTileBinEdges", "MapGaussiansToIntersects",
                 "SphericalHarmonics", "NDRasterizeGaussians"):
        assert hasattr(rasterizer, name), name
    # the 11 native bindings of csrc/ext.cpp:6-17
    for name in ("nd_rasterize_forward", "nd_rasterize_backward", "rasterize_forward", "rasterize_backward",
                 "project_gaussians_forward", "project_gaussians_backward", "compute_sh_forward",
                 "compute_sh_backward", "compute_cov2d_bounds", "map_gaussian_to_intersects", "get_tile_bin_edges"):
        assert callable(getattr(C, name)), name
    # argument names / order of the three operators (rasterizer/project_gaussians.py:12-27, rasterize.py:14-27,
    # sh.py:36-40)
    assert list(inspect.signature(pg_mod.project_gaussians).parameters) == [
        "means3d", "scales", "glob_scale", "quats", "viewmat", "projmat", "fx", "fy", "cx", "cy", "img_height",
        "img_width", "block_width", "clip_thresh"]
    assert list(inspect.signature(rz_mod.rasterize_gaussians).parameters) == [
        "xys", "depths", "radii", "conics", "num_tiles_hit", "colors", "opacity", "img_height", "img_width",
        "block_width", "background", "return_alpha"]
    assert list(inspect.signature(sh_mod.spherical_harmonics).parameters) == ["degrees_to_use", "viewdirs", "coeffs"]
    assert [sh_mod.num_sh_bases(d) for d in range(6)] == [1, 4, 9, 16, 25, 25]
    assert [sh_mod.deg_from_sh(k) for k in (1, 4, 9, 16, 25)] == [0, 1, 2, 3, 4]
    q = torch.tensor([[2.0, 0.0, 0.0, 0.0], [0.0, 0.0, 0.0, 3.0]])
    R = quat_to_rotmat(q)
    assert torch.allclose(R[0], torch.eye(3)) and torch.allclose(R[1], torch.diag(torch.tensor([-1.0, -1.0, 1.0])))


def test_python_error_behaviour_on_cpu():
    import rasterizer

    xys = torch.zeros(4, 2)
    with pytest.raises(AssertionError, match="block_width"):
        rasterizer.rasterize_gaussians(xys, torch.zeros(4), torch.zeros(4, dtype=torch.int32), torch.zeros(4, 3),
                                       torch.zeros(4, dtype=torch.int32), torch.zeros(4, 3), torch.zeros(4, 1),
                                       16, 16, 17)
    with pytest.raises(ValueError, match="xys must have dimensions"):
        rasterizer.rasterize_gaussians(torch.zeros(4, 3), torch.zeros(4), torch.zeros(4, dtype=torch.int32),
                                       torch.zeros(4, 3), torch.zeros(4, dtype=torch.int32), torch.zeros(4, 3),
                                       torch.zeros(4, 1), 16, 16, 16)
    with pytest.raises(AssertionError, match="background"):
        rasterizer.rasterize_gaussians(xys, torch.zeros(4), torch.zeros(4, dtype=torch.int32), torch.zeros(4, 3),
                                       torch.zeros(4, dtype=torch.int32), torch.zeros(4, 3), torch.zeros(4, 1),
                                       16, 16, 16, background=torch.zeros(4))
    with pytest.raises(AssertionError, match="block_width"):
        rasterizer.project_gaussians(torch.zeros(4, 3), torch.zeros(4, 3), 1.0, torch.zeros(4, 4), torch.eye(4),
                                     torch.eye(4), 1.0, 1.0, 0.0, 0.0, 16, 16, 1)
    # CPU tensors are rejected like CHECK_CUDA does (bindings.h:10) — no silent CPU path
    with pytest.raises(RuntimeError, match="must be a CUDA tensor"):
        rasterizer.spherical_harmonics(0, torch.zeros(4, 3), torch.zeros(4, 1, 3))
    with pytest.raises(RuntimeError, match="must be a CUDA tensor"):
        rasterizer.project_gaussians(torch.zeros(4, 3), torch.zeros(4, 3), 1.0, torch.zeros(4, 4), torch.eye(4),
                                     torch.eye(4), 1.0, 1.0, 0.0, 0.0, 16, 16, 16)
    with pytest.warns(DeprecationWarning):
        with pytest.raises(RuntimeError):
            rasterizer.SphericalHarmonics.apply(0, torch.zeros(4, 3), torch.zeros(4, 1, 3))


def test_tile_mask_division_magic_is_exact():
    """binning_device.cu walks the cached 64-bit tile mask of a bounding box without an integer division: row of bit k
    = (k * magic) >> 16 with magic = trunc(65536 / d) + 1 from a fast-math FP32 reciprocal (d = box width in tiles, at
    most 64; k < 64).  Every magic the approximate division can produce must give exactly k // d."""
    for d in range(1, 65):
        x = 65536.0 / d
        for magic in {int(x * (1 - 2.0 ** -20)) + 1, int(x) + 1, int(x * (1 + 2.0 ** -20)) + 1}:
            for k in range(64):
                assert (k * magic) >> 16 == k // d, (d, magic, k)


def test_nvtx_tracing_switch_is_harmless_without_a_tool():
    """GSR_NVTX=1 wraps every C-ABI entry point in an NVTX range and marks every kernel launch (csrc/api.cu); without a
    profiler attached the NVTX 3 calls are no-ops.  Runs in its own interpreter: the switch is read once per process."""
    import subprocess
    import sys

    code = (
        "import sys, os\n"
        f"sys.path.insert(0, os.path.join({ROOT!r}, 'gaussian-splatting-toolkit_b200'))\n"
        "from rasterizer import _lib\n"
        "lib = _lib.load()\n"
        "rc = lib.gsr_rasterize_forward(8, 8, 1, 0, None, None, None, None, None, None, None, None, None, None, None)\n"
        "assert rc != 0 and b'block_width' in lib.gsr_last_error(), (rc, lib.gsr_last_error())\n"
        "print('ok')\n"
    )
    out = subprocess.run([sys.executable, "-c", code], env=dict(os.environ, GSR_NVTX="1"), capture_output=True, text=True)
    assert out.returncode == 0 and "ok" in out.stdout, out.stderr


def test_mask_bit_gather_multiply_is_exact():
    """compact_from_masks (csrc/blend_common.cuh) collects bit `warp` of eight mask bytes with two multiplies:
    ((x >> warp) & 0x01010101) * 0x01020408 puts the four byte flags into bits 24..27 without carries."""
    import itertools

    for warp in range(8):
        for flags in itertools.product((0, 1), repeat=4):
            for noise in (0x00, 0xFF, 0xA5):  # the other bits of the bytes must not matter
                word = 0
                for i, f in enumerate(flags):
                    byte = (noise & ~(1 << warp) & 0xFF) | (f << warp)
                    word |= byte << (8 * i)
                lo = (word >> warp) & 0x01010101
                got = ((lo * 0x01020408) & 0xFFFFFFFF) >> 24
                assert got & 0xF == sum(f << i for i, f in enumerate(flags)), (warp, flags, noise)
                assert got >> 4 == 0
