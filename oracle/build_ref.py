"""TEST INFRASTRUCTURE — builds the UNMODIFIED reference CUDA extension into oracle/_ref/.

The reference rasterizer (gs_toolkit/gs_components/rasterizer/cuda/csrc/{ext.cpp,bindings.cu,
forward.cu,backward.cu} + vendored glm) is compiled *from the sources where they lie* under
/root/reference; nothing is copied into this repository.  Only the resulting shared object
``oracle/_ref/rasterizer_ref_cuda.so`` (git-ignored, NOT gpurun-ignored) travels to the GPU box, where
the parity tests and the reference-timing test load it with :func:`load_ref`.

Flags follow the reference's packaged build (gs_toolkit/gs_components/setup.py:73-85:
``-O3 --use_fast_math --expt-relaxed-constexpr``); the arch is sm_100 (what
``TORCH_CUDA_ARCH_LIST=10.0`` would give the reference's own setup.py).

Only tests/, __graft_entry__.smoke()/build() and bench.py's reference legs may use this module.
"""
from __future__ import annotations

import glob
import importlib.util
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
OUT_DIR = os.path.join(HERE, "_ref")
NAME = "rasterizer_ref_cuda"
REF_CSRC = "/root/reference/gs_toolkit/gs_components/rasterizer/cuda/csrc"


def ref_so_path() -> str | None:
    hits = glob.glob(os.path.join(OUT_DIR, NAME + "*.so"))
    return hits[0] if hits else None


def build_ref(verbose: bool = False) -> str | None:
    """Compile the reference extension if its sources are present; return the .so path."""
    so = ref_so_path()
    if so is not None:
        return so
    if not os.path.isdir(REF_CSRC):
        return None
    os.makedirs(OUT_DIR, exist_ok=True)
    os.environ.setdefault("TORCH_CUDA_ARCH_LIST", "10.0")
    os.environ.setdefault("MAX_JOBS", "8")
    from torch.utils.cpp_extension import load

    sources = sorted(glob.glob(os.path.join(REF_CSRC, "*.cu"))) + sorted(
        glob.glob(os.path.join(REF_CSRC, "*.cpp"))
    )
    load(
        name=NAME,
        sources=sources,
        extra_include_paths=[os.path.join(REF_CSRC, "third_party", "glm")],
        extra_cflags=["-O3"],
        extra_cuda_cflags=["-O3", "--use_fast_math", "--expt-relaxed-constexpr"],
        build_directory=OUT_DIR,
        is_python_module=False,  # only build; load_ref() imports it explicitly
        verbose=verbose,
    )
    # keep only the shared object (no objects / ninja files: nothing derived from sources but the .so)
    for f in os.listdir(OUT_DIR):
        if not f.endswith(".so"):
            try:
                os.remove(os.path.join(OUT_DIR, f))
            except OSError:
                pass
    return ref_so_path()


def load_ref():
    """Import the prebuilt reference extension (pybind module exposing the 11 reference bindings)."""
    so = ref_so_path()
    if so is None:
        return None
    import torch  # noqa: F401  (libtorch symbols must be loaded first)

    if NAME in sys.modules:
        return sys.modules[NAME]
    spec = importlib.util.spec_from_file_location(NAME, so)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    sys.modules[NAME] = mod
    return mod


if __name__ == "__main__":
    p = build_ref(verbose=True)
    print("reference extension:", p)
