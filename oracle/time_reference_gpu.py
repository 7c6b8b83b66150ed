#!/usr/bin/env python3
"""TEST / BENCH INFRASTRUCTURE ONLY.  Times the UNMODIFIED compiled reference (oracle/_ref, built by `make -C oracle ref`)
on the GPU through its own front end (PASIf.py) on a bounded sample of a bench workload, in its own process (the
reference calls std::exit(0) on CUDA errors, src/helpers.cuh:43-51).  Prints one JSON line.

    python oracle/time_reference_gpu.py --prec f32 --systems 2 --excitations 64 --steps 20000

Limits of the reference that bound the sample: its front end cannot assemble more than 2 Jacobians (PASIf.py:574), and
its Lambda kernel is only correct while nnz(Gamma)+nnz(Lambda)+nnz(Psi) <= 128."""
import argparse
import json
import os
import sys
import time

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
REF = os.path.join(HERE, "_ref")


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--prec", default="f64")
    ap.add_argument("--systems", type=int, default=2)
    ap.add_argument("--excitations", type=int, default=64)
    ap.add_argument("--steps", type=int, default=20000)
    ap.add_argument("--exc-seed", type=int, default=2)
    ap.add_argument("--forward-only", action="store_true",
                    help="no Jacobian / gradient: the only way to time the reference on more than 2 systems (its front end cannot "
                         "assemble more Jacobians); above nnz 128 its results are invalid (quirk F4), its time is still its time")
    a = ap.parse_args()
    sys.path.insert(0, os.path.join(REF, a.prec))
    sys.path.insert(0, REF)
    sys.path.insert(0, ROOT)
    import numpy as np
    import PASIf as ref
    from pasif_b200 import workloads as wl
    rng = np.random.default_rng(1)
    systems = [wl.jitter(wl.fit8_system(), rng) for _ in range(a.systems)]
    jacs = [wl.fit8_jacobian() for _ in range(a.systems)]
    exc = [np.array(r) for r in wl.random_sine_excitations(a.excitations, a.steps, 16000, seed=a.exc_seed)]
    stdout = os.dup(1)
    os.dup2(2, 1)                    # the reference prints "[Info] ..." lines on stdout
    p = ref.PASIf(exc, 16000, 0)
    t0 = time.time()
    p.setSystems(*wl.to_frontend_args(systems, ref))
    t_set = time.time() - t0
    p.getAmplitudes()                # warm-up (graph capture, allocations)
    t0 = time.time()
    p.getAmplitudes()
    t_f = time.time() - t0
    if a.forward_only:
        sys.stdout.flush()
        os.dup2(stdout, 1)
        print(json.dumps({"prec": a.prec, "systems": a.systems, "excitations": a.excitations, "steps": a.steps, "forward_s": t_f,
                          "setSystems_s": t_set, "forward_system_steps_per_s": a.systems * a.excitations * a.steps / t_f,
                          "results_valid": bool(a.systems * 5 <= 128)}))
        return
    p.setJacobian(*wl.to_frontend_args(jacs, ref, with_psi=True))
    t0 = time.time()
    p.getGradient()
    t_g = time.time() - t0
    sys.stdout.flush()
    os.dup2(stdout, 1)
    pairs = a.systems * a.excitations
    print(json.dumps({"prec": a.prec, "systems": a.systems, "excitations": a.excitations, "steps": a.steps,
                      "forward_s": t_f, "gradient_s": t_g,
                      "forward_system_steps_per_s": pairs * a.steps / t_f,
                      "gradient_trajectory_steps_per_s": pairs * a.steps / t_g}))


if __name__ == "__main__":
    main()
