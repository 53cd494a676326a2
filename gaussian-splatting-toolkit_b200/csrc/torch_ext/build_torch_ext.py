"""Builds `rasterizer/csrc*.so` — the pybind module with the reference's eleven binding names (ext.cpp) — in-tree with
g++ (host code only; the kernels live in libgsr_b200.so, found at run time through an $ORIGIN-relative rpath)."""
from __future__ import annotations

import os
import subprocess
import sysconfig

HERE = os.path.dirname(os.path.abspath(__file__))
PKG = os.path.dirname(os.path.dirname(HERE))  # gaussian-splatting-toolkit_b200/
OUT = os.path.join(PKG, "rasterizer", "csrc.so")
SRC = os.path.join(HERE, "ext.cpp")


def build(force: bool = False, verbose: bool = False) -> str:
    lib = os.path.join(PKG, "libgsr_b200.so")
    if not os.path.exists(lib):
        raise RuntimeError("build libgsr_b200.so first (make -C csrc)")
    hdr = os.path.join(os.path.dirname(PKG), "include", "gsr_b200.h")
    if not force and os.path.exists(OUT) and os.path.getmtime(OUT) >= max(os.path.getmtime(p) for p in (SRC, hdr)):
        return OUT
    import torch
    from torch.utils import cpp_extension as ce

    inc = ce.include_paths("cuda") + [sysconfig.get_paths()["include"]]
    libdirs = ce.library_paths("cuda")
    cmd = ["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-Wno-sign-compare", SRC, "-o", OUT,
           "-DTORCH_EXTENSION_NAME=csrc", "-DTORCH_API_INCLUDE_EXTENSION_H",
           f"-D_GLIBCXX_USE_CXX11_ABI={int(torch._C._GLIBCXX_USE_CXX11_ABI)}"]
    cmd += [f"-I{p}" for p in inc] + [f"-L{p}" for p in libdirs] + [f"-L{PKG}"]
    cmd += ["-lgsr_b200", "-lc10", "-lc10_cuda", "-ltorch_cpu", "-ltorch_cuda", "-ltorch", "-ltorch_python", "-lcudart",
            "-Wl,-rpath,$ORIGIN/..", "-Wl,--no-as-needed"]
    cmd += [f"-Wl,-rpath,{p}" for p in libdirs]
    if verbose:
        print(" ".join(cmd))
    subprocess.check_call(cmd)
    return OUT


if __name__ == "__main__":
    print(build(force=True, verbose=True))
