import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from pasif_b200 import capi
for N,P in ((128,512),(256,1024),(128,256)):
    rng=np.random.default_rng(N+P)
    kd, x = rng.standard_normal((N,N)), rng.standard_normal((P,N))
    ref = x @ kd.T
    y2 = capi.debug_gemm(kd, x, "f32", 2)
    y3 = capi.debug_gemm(kd, x, "f32", 3)
    print(N,P, np.max(np.abs(y2-ref))/np.max(np.abs(ref)), np.max(np.abs(y3-ref))/np.max(np.abs(ref)), np.array_equal(y2,y3), flush=True)
