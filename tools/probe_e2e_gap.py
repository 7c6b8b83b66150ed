#!/usr/bin/env python3
"""Device time of a C3 gradient call with resident excitations (G) against a call that follows a fresh upload (L then G)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from pasif_b200 import capi, frontend, workloads as wl


class mod:
    coo_tensor = frontend.coo_tensor


rng = np.random.default_rng(1)
systems = [wl.jitter(wl.fit8_system(), rng) for _ in range(4096)]
jacs = [wl.fit8_jacobian() for _ in range(4096)]
exc = np.ascontiguousarray(wl.random_sine_excitations(64, 100000, 16000, seed=2))
p = frontend.make_frontend(capi.engine_class("f32"))(exc, 16000, 0)
for kv in sys.argv[1:]:
    k, v = kv.split("=")
    p._setOption(k, v)
p.setSystems(*wl.to_frontend_args(systems, mod))
p.setJacobian(*wl.to_frontend_args(jacs, mod, with_psi=True))
out = []
for c in "GGGGLGLGLGGGFGFGLGLGGG":
    if c == "G":
        p._getGradient(0, fetch=False)
        out.append(("G", round(p._query("last_device_ms"), 1)))
    elif c == "F":
        p._getGradient(0)
        out.append(("G+fetch", round(p._query("last_device_ms"), 1)))
    else:
        p._loadExcitationsSet(exc, 16000)
        out.append(("L",))
print(out)
