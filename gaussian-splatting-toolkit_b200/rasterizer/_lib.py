"""ctypes loader for libgsr_b200.so — the C ABI declared in include/gsr_b200.h.

There is NO fallback: if the shared library is missing or a call fails, a RuntimeError / ImportError is
raised.  (The reference loader tries a prebuilt `rasterizer.csrc`, then JIT-compiles, and prints
"gsplat: No CUDA toolkit found" otherwise — rasterizer/cuda/_backend.py:61-100.)
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
import threading

_PKG_DIR = os.path.dirname(os.path.abspath(__file__))
_ROOT = os.path.dirname(_PKG_DIR)  # gaussian-splatting-toolkit_b200/
LIB_PATH = os.path.join(_ROOT, "libgsr_b200.so")
CSRC_DIR = os.path.join(_ROOT, "csrc")

_lock = threading.Lock()
_lib = None

_i, _u, _f, _d, _p, _sz = C.c_int, C.c_uint, C.c_float, C.c_double, C.c_void_p, C.c_size_t

# name -> (restype, argtypes); kept in the order of include/gsr_b200.h
SIGNATURES = {
    "gsr_version": (C.c_char_p, []),
    "gsr_last_error": (C.c_char_p, []),
    "gsr_built_for_sm": (_i, []),
    "gsr_launch_count": (C.c_ulonglong, []),
    "gsr_compute_sh_forward": (_i, [_i, _i, _i, _p, _p, _p, _p]),
    "gsr_compute_sh_backward": (_i, [_i, _i, _i, _p, _p, _p, _p]),
    "gsr_compute_sh_backward_multiview": (_i, [_i, _i, _i, _i, _p, _p, _p, _p, _p]),
    "gsr_compute_sh_backward_multiview_ptrs": (_i, [_i, _i, _i, _i, _p, _p, _p, _p, _p]),
    "gsr_peer_reduce_scatter": (_i, [_i, _i, _p, C.c_longlong, _p]),
    "gsr_peer_all_gather": (_i, [_i, _i, _p, C.c_longlong, _p]),
    "gsr_peer_push": (_i, [_i, _p, _p, C.c_longlong, C.c_longlong, C.c_longlong, _p]),
    "gsr_peer_reduce_broadcast": (_i, [_i, _p, _p, C.c_longlong, C.c_longlong, _p]),
    "gsr_project_gaussians_forward": (_i, [_i, _p, _p, _f, _p, _p, _p, _f, _f, _f, _f, _u, _u, _u, _f,
                                           _p, _p, _p, _p, _p, _p, _p, _p]),
    "gsr_project_gaussians_backward": (_i, [_i, _p, _p, _f, _p, _p, _p, _f, _f, _f, _f, _u, _u, _p, _p, _p,
                                            _p, _p, _p, _p, _p, _p, _p, _p, _p, _p, _p]),
    "gsr_compute_cov2d_bounds": (_i, [_i, _p, _p, _p, _p]),
    "gsr_cumsum_workspace_bytes": (_sz, [_i]),
    "gsr_cumsum_tiles_hit": (_i, [_i, _p, _p, _p, _p, _sz, _p]),
    "gsr_map_gaussian_to_intersects": (_i, [_i, _i, _p, _p, _p, _p, _u, _u, _u, _p, _p, _p]),
    "gsr_count_tiles_tight": (_i, [_i, _p, _p, _p, _p, _u, _u, _u, _p, _p]),
    "gsr_map_gaussian_to_intersects_tight": (_i, [_i, _i, _p, _p, _p, _p, _p, _p, _u, _u, _u, _p, _p, _p]),
    "gsr_bin_device_workspace_bytes": (_sz, [_i, _i, _u, _u, _u]),
    "gsr_bin_gaussians_device": (_i, [_i, _p, _p, _p, _p, _p, _u, _u, _u, _i, _p, _p, _p, _p, _p, _sz, _p]),
    "gsr_bin_count_workspace_bytes": (_sz, [_i, _u, _u, _u]),
    "gsr_bin_count": (_i, [_i, _p, _p, _p, _p, _u, _u, _u, _p, _p, _p, _p, _sz, _p]),
    "gsr_bin_fill_workspace_bytes": (_sz, [_i]),
    "gsr_bin_fill_sort": (_i, [_i, _i, _p, _p, _p, _p, _p, _u, _u, _u, _p, _p, _p, _p, _sz, _p]),
    "gsr_sort_workspace_bytes": (_sz, [_i]),
    "gsr_sort_intersects": (_i, [_i, _i, _p, _p, _p, _p, _p, _sz, _p]),
    "gsr_get_tile_bin_edges": (_i, [_i, _p, _i, _p, _p]),
    "gsr_rasterize_forward": (_i, [_u, _u, _u, _i, _p, _p, _p, _p, _p, _p, _p, _p, _p, _p, _p]),
    "gsr_rasterize_backward": (_i, [_u, _u, _u, _i, _p, _p, _p, _p, _p, _p, _p, _p, _p, _p, _p, _p, _p,
                                    _p, _p, _p]),
    "gsr_nd_rasterize_forward": (_i, [_u, _u, _u, _u, _i, _p, _p, _p, _p, _p, _p, _p, _p, _p, _p, _p]),
    "gsr_nd_rasterize_backward": (_i, [_u, _u, _u, _u, _i, _p, _p, _p, _p, _p, _p, _p, _p, _p, _p, _p, _p,
                                       _p, _p, _p, _p]),
    "gsr_fused_preprocess_forward": (_i, [_i, _i, _i, _p, _p, _p, _p, _p, _p, _p, _p, _f, _f, _f, _f, _f, _u, _u, _u, _f,
                                          _p, _p, _p, _p, _p, _p, _p, _p, _p]),
    "gsr_fused_preprocess_backward": (_i, [_i, _i, _i, _p, _p, _p, _p, _p, _p, _f, _f, _f, _u, _u, _p, _p, _p, _p, _p, _p,
                                           _p, _p, _p, _p, _p, _p, _p]),
    "gsr_blend_packed_forward": (_i, [_u, _u, _u, _i, _p, _p, _p, _p, _p, _p, _p, _p, _p]),
    "gsr_blend_packed_backward": (_i, [_u, _u, _u, _i, _p, _p, _p, _p, _p, _p, _p, _p, _p, _p, _p]),
    "gsr_l1_ssim_num_partials": (_i, [_u, _u]),
    "gsr_l1_ssim_forward": (_i, [_u, _u, _f, _p, _p, _p, _p, _p, _p]),
    "gsr_l1_ssim_backward": (_i, [_u, _u, _f, _p, _p, _p, _p, _p, _p]),
    "gsr_adam_step_multi": (_i, [_i, _p, _p, _p, _p, _p, _p, _p, _d, _d, _d, _f, _p]),
    "gsr_opacity_reset": (_i, [_i, _f, _p, _p, _p, _p]),
    "gsr_densify_stats_update": (_i, [_i, _p, _i, _p, _f, _i, _p, _p, _p, _p]),
    "gsr_densify_plan_workspace_bytes": (_sz, [_i]),
    "gsr_densify_plan": (_i, [_i, _p, _p, _p, _p, _p, _i, _f, _f, _f, _i, _f, _f, _i, _f, _i, _f, _p, _p, _p, _p, _sz, _p]),
    "gsr_densify_apply": (_i, [_i, _i, _p, _p, _p, _p, _p, _p, _p, _i, _p, _p, _p, _p, _p, _p]),
}


def build_library(verbose: bool = False) -> str:
    """Compile libgsr_b200.so in-tree with nvcc for sm_100a (see csrc/Makefile)."""
    cmd = ["make", "-C", CSRC_DIR, "-j", str(os.cpu_count() or 4)]
    res = subprocess.run(cmd, capture_output=not verbose, text=True)
    if res.returncode != 0:
        raise RuntimeError("building libgsr_b200.so failed:\n" + (res.stdout or "") + (res.stderr or ""))
    return LIB_PATH


def load():
    """Load the shared library once and attach the prototypes."""
    global _lib
    if _lib is not None:
        return _lib
    with _lock:
        if _lib is not None:
            return _lib
        if not os.path.exists(LIB_PATH):
            raise ImportError(
                f"{LIB_PATH} not found. The B200 rasterizer has no CPU / PyTorch fallback: build it with "
                f"`python -c 'import __graft_entry__ as g; g.build()'` or `make -C {CSRC_DIR}`.")
        lib = C.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(lib, name)  # AttributeError => the .so does not match include/gsr_b200.h
            fn.restype = res
            fn.argtypes = args
        _lib = lib
    return _lib


def check(rc: int, what: str = "") -> None:
    if rc != 0:
        msg = load().gsr_last_error().decode("utf-8", "replace")
        raise RuntimeError(f"libgsr_b200 {what} failed (status {rc}): {msg}")
