#!/usr/bin/env python3
"""A/B of the two-trajectories-per-thread fp32 kernels (packed FFMA2, option pair=on) against the one-trajectory kernels
on the C3 workload at a reduced step count: bit-identity of amplitudes and gradient, device times, FMA-pipe peaks.

    python tools/probe_pair.py [--systems 4096] [--excitations 64] [--steps 20000] [--option k=v ...]
"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
from pasif_b200 import capi, frontend, workloads as wl  # noqa: E402


class mod:
    coo_tensor = frontend.coo_tensor


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--systems", type=int, default=4096)
    ap.add_argument("--excitations", type=int, default=64)
    ap.add_argument("--steps", type=int, default=20000)
    ap.add_argument("--reps", type=int, default=3)
    ap.add_argument("--compat", default="reference")
    ap.add_argument("--option", action="append", default=[])
    ap.add_argument("--variants", default="pair=off;pair=on",
                    help="';'-separated engine option sets (each 'k=v,k=v'); the first is the bit-identity reference")
    a = ap.parse_args()
    rng = np.random.default_rng(1)
    systems = [wl.jitter(wl.fit8_system(), rng) for _ in range(a.systems)]
    jacs = [wl.fit8_jacobian() for _ in range(a.systems)]
    exc = np.ascontiguousarray(wl.random_sine_excitations(a.excitations, a.steps, 16000, seed=2))
    out = {"systems": a.systems, "excitations": a.excitations, "steps": a.steps, "compat": a.compat}
    res = {}
    variants = [v for v in a.variants.split(";") if v]
    for var in variants:
        p = frontend.make_frontend(capi.engine_class("f32"))(exc, 16000, 0)
        p._setOption("compat", a.compat)
        for kv in a.option + var.split(","):
            k, v = kv.split("=", 1)
            p._setOption(k, v)
        p.setSystems(*wl.to_frontend_args(systems, mod))
        p.setJacobian(*wl.to_frontend_args(jacs, mod, with_psi=True))
        amp = np.asarray(p._getAmplitudes())
        grad = np.asarray(p._getGradient(0))
        fwd_ms, grad_ms = [], []
        for _ in range(a.reps):
            p._getAmplitudes(fetch=False)
            fwd_ms.append(p._query("last_device_ms"))
            p._getGradient(0, fetch=False)
            grad_ms.append(p._query("last_device_ms"))
        p._setOption("profile", "on")
        p._getGradient(0, fetch=False)
        k = {n: p._query("ms_" + n) for n in ("gfwd", "bwd")}
        res[var] = (amp, grad)
        pairs = a.systems * a.excitations
        out[var] = {"bit_identical_to_first": bool(np.array_equal(amp, res[variants[0]][0]) and np.array_equal(grad, res[variants[0]][1])),
                    "fwd_ms": min(fwd_ms), "grad_ms": min(grad_ms), "grad_evals_per_s_at_100k": pairs / (min(grad_ms) * 1e-3) * a.steps / 100000,
                               "fwd_tflops": pairs * a.steps * 240.0 / (min(fwd_ms) * 1e-3) / 1e12,
                               "serial_gfwd_ms": k["gfwd"], "serial_bwd_ms": k["bwd"], "pair_launches": p._query("pair_launches")}
        p.close()
    out["fma_peak_tflops"] = {k: capi.fma_peak(0, k) for k in ("f32", "f32x2", "f64")}
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
