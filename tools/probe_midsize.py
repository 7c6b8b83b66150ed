#!/usr/bin/env python3
"""Mid-size components (13..48 coupled variables): one thread per trajectory against the lane-block split.
    python tools/probe_midsize.py [ndof ...]"""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from pasif_b200 import capi, frontend, workloads as wl


class mod:
    coo_tensor = frontend.coo_tensor


def run(ndof, prec, opts, E=65536, L=600):
    s = wl.chain_system(ndof)
    exc = wl.random_sine_excitations(E, L, 16000, seed=1)
    P = frontend.make_frontend(capi.engine_class(prec))
    p = P(exc, 16000, 0)
    for k, v in opts.items():
        p._setOption(k, v)
    p.setSystems(*wl.to_frontend_args([s], mod))
    p.setJacobian(*wl.to_frontend_args([wl.chain_jacobian(s, 3)], mod, with_psi=True))
    for _ in range(2):
        p._getAmplitudes(fetch=False)
    f = p._query("last_device_ms")
    for _ in range(2):
        p._getGradient(0, fetch=False)
    g = p._query("last_device_ms")
    return dict(fwd_steps_per_s=E * L / f * 1e3, grad_ms=g, lanes=p._query("lane_classes"), G=p._query("lane_G"))


if __name__ == "__main__":
    for nd in [int(x) for x in sys.argv[1:]] or [6, 10, 16, 23]:
        for prec in ("f32", "f64"):
            a = run(nd, prec, {"one_thread_regs": "100000"})
            b = run(nd, prec, {"one_thread_regs": "0"})
            print(nd, prec, "one-thread", a, "| lanes", b, "| speedup fwd", b["fwd_steps_per_s"] / a["fwd_steps_per_s"], "grad", a["grad_ms"] / b["grad_ms"], flush=True)
