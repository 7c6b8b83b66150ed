#!/usr/bin/env python3
"""Isolated check of the dense-operator path's GEMM kernels against numpy (python tools/probe_gemm.py [N] [P])."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from pasif_b200 import capi

N = int(sys.argv[1]) if len(sys.argv) > 1 else 100
P = int(sys.argv[2]) if len(sys.argv) > 2 else 300
rng = np.random.default_rng(0)
kd = rng.standard_normal((N, N))
x = rng.standard_normal((P, N))
ref = x @ kd.T
for dtype, mode, name in (("f64", 0, "DMMA"), ("f32", 0, "3xTF32 mma.sync"), ("f32", 1, "3xTF32 tcgen05")):
    y = capi.debug_gemm(kd, x, dtype, mode)
    err = np.max(np.abs(y - ref)) / np.max(np.abs(ref))
    print(f"{name:18s} N={N} P={P} max rel err {err:.3e}", flush=True)
