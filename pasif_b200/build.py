"""In-tree build of the native pieces (explicit nvcc / g++ commands, no JIT cache outside the repo).

    libpasif_b200.so                    C-ABI engine + CUDA kernels  (csrc/engine.cu, sm_100a)
    PASIfgpu.cpython-*.so               pybind11 module `PASIfgpu.__GpuDriver` over the C ABI (csrc/PASIfgpu.cpp)

Both land in ``pasif_b200/`` so they travel with the repository snapshot to the GPU box.
"""
from __future__ import annotations

import os
import subprocess
import sys
import sysconfig

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libpasif_b200.so")
MOD = os.path.join(HERE, "PASIfgpu" + (sysconfig.get_config_var("EXT_SUFFIX") or ".so"))

NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
              "-Xcompiler", "-fPIC", "-shared", "-cudart", "static"]


def _newer(target, sources):
    if not os.path.exists(target):
        return False
    t = os.path.getmtime(target)
    return all(os.path.getmtime(s) <= t for s in sources if os.path.exists(s))


def _run(cmd):
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if r.returncode != 0:
        raise RuntimeError("build command failed:\n  " + " ".join(cmd) + "\n" + r.stdout)
    return r.stdout


def build_library(force=False, verbose=False):
    srcs = [os.path.join(CSRC, f) for f in ("engine.cu", "rk4_kernels.cuh", "large_kernels.cuh", "large_host.inc", "stage_kernels.cuh", "stage_tcgen05.cuh", "stage_host.inc", "pattern.hpp", "lane_split.hpp", "patterns_aot.inc")] + \
           [os.path.join(os.path.dirname(HERE), "include", "pasif_b200.h")]
    if not force and _newer(LIB, srcs):
        return LIB
    cmd = ["nvcc"] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-o", LIB, os.path.join(CSRC, "engine.cu"), "-ldl", "-lpthread"]
    out = _run(cmd)
    if verbose:
        print(out)
    return LIB


def build_module(force=False):
    import pybind11
    srcs = [os.path.join(CSRC, "PASIfgpu.cpp"), os.path.join(os.path.dirname(HERE), "include", "pasif_b200.h"), LIB]
    if not force and _newer(MOD, srcs):
        return MOD
    cmd = ["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-fvisibility=hidden",
           "-I" + sysconfig.get_paths()["include"], "-I" + pybind11.get_include(),
           "-o", MOD, os.path.join(CSRC, "PASIfgpu.cpp"),
           "-L" + HERE, "-lpasif_b200", "-Wl,-rpath,$ORIGIN"]
    _run(cmd)
    return MOD


def build_all(force=False, verbose=False):
    build_library(force, verbose)
    build_module(force)
    return LIB, MOD


if __name__ == "__main__":
    build_all(force="--force" in sys.argv, verbose="-v" in sys.argv)
    print(LIB)
    print(MOD)
