import sys; sys.path.insert(0,'/root/repo')
from pasif_b200 import capi
for k in ("f64","f32","dmma","tf32_mma"): print(k, capi.fma_peak(0,k))
