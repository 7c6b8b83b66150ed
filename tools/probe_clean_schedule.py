import sys, os, time
sys.path.insert(0, '/root/repo')
import numpy as np
from pasif_b200 import capi, frontend, workloads as wl
class mod: coo_tensor = frontend.coo_tensor
prec = sys.argv[1]
s = wl.chain_system(256)
exc = wl.random_sine_excitations(16384, 2000, 16000, seed=5)
P = frontend.make_frontend(capi.engine_class(prec))
p = P(exc, 16000, 0)
p._setOption("compat", "clean")
for kv in sys.argv[2:]:
    k, v = kv.split("="); p._setOption(k, v)
p.setSystems(*wl.to_frontend_args([s], mod))
p.setJacobian(*wl.to_frontend_args([wl.chain_jacobian(s, 4)], mod, with_psi=True))
seq = "GGGGAGAGGGG"
out = []
for c in seq:
    t = time.time()
    if c == "G": p._getGradient(0, fetch=False)
    else: p._getAmplitudes(fetch=False)
    out.append((c, round(p._query("last_device_ms"), 1), round((time.time() - t) * 1e3, 1), int(p._query("stored_chunks"))))
print(prec, sys.argv[2:], out)
