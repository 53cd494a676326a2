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


# --------------------------------------------------------------------------------------------------------------------
# The reference's PYTHON side under baseline/_ref/ (git-ignored, travels to the GPU box) — used by bench.py's literal
# `_torch_impl` timing and by tests/test_gpu_dropin_reference_callers.py (the reference's unmodified wrappers / model
# class driven over libgsr_b200).  Two steps, both only where /root/reference exists:
#   1. `pip install --no-index --no-build-isolation --no-deps --target baseline/_ref <copy of gs_toolkit/gs_components>`
#      — the reference's own setup.py (package `rasterizer` + its CUDA extension `rasterizer/csrc.so`, sm_100);
#   2. the pure-Python `gs_toolkit` package: its build backend (poetry-core) is absent from this image, so
#      `pip install /root/reference` cannot run; the package directory is copied as pip would have copied it
#      (minus gs_components — step 1 — and the legacy viewer's static assets).
BASELINE_REF = os.path.join(os.path.dirname(HERE), "baseline", "_ref")
REF_ROOT = "/root/reference"


def install_ref_python(verbose: bool = False) -> str | None:
    import shutil
    import subprocess
    import tempfile

    if not os.path.isdir(os.path.join(REF_ROOT, "gs_toolkit")):
        return BASELINE_REF if os.path.isdir(os.path.join(BASELINE_REF, "rasterizer")) else None
    os.makedirs(BASELINE_REF, exist_ok=True)
    if not os.path.exists(os.path.join(BASELINE_REF, "rasterizer", "csrc.so")):
        with tempfile.TemporaryDirectory() as tmp:
            src = os.path.join(tmp, "gs_components")
            shutil.copytree(os.path.join(REF_ROOT, "gs_toolkit", "gs_components"), src)
            env = dict(os.environ, TORCH_CUDA_ARCH_LIST="10.0", MAX_JOBS="8")
            subprocess.run([sys.executable, "-m", "pip", "install", "--no-index", "--no-build-isolation", "--no-deps",
                            "--find-links", "/opt/wheelhouse", "--target", BASELINE_REF, src], check=True, env=env,
                           stdout=None if verbose else subprocess.DEVNULL, stderr=None if verbose else subprocess.DEVNULL)
    dst = os.path.join(BASELINE_REF, "gs_toolkit")
    if not os.path.isdir(dst):
        shutil.copytree(os.path.join(REF_ROOT, "gs_toolkit"), dst,
                        ignore=shutil.ignore_patterns("gs_components", "viewer_legacy", "__pycache__", "*.so"))
    return BASELINE_REF
