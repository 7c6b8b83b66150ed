#!/usr/bin/env python3
"""256-row-tile variant of the fused-split tcgen05 GEMM (option tc_bn=256) against numpy and against the 128-row tiles."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from pasif_b200 import capi

for N, P in ((256, 512), (512, 1024), (256, 300)):
    rng = np.random.default_rng(N + P)
    kd, x = rng.standard_normal((N, N)), rng.standard_normal((P, N))
    ref = x @ kd.T
    y2 = capi.debug_gemm(kd, x, "f32", 2)
    y4 = capi.debug_gemm(kd, x, "f32", 4)
    print(N, P, np.max(np.abs(y2 - ref)) / np.max(np.abs(ref)), np.max(np.abs(y4 - ref)) / np.max(np.abs(ref)), np.array_equal(y2, y4), flush=True)
