#!/usr/bin/env python3
"""Generate golden vectors by running the UNMODIFIED compiled reference (oracle/_ref, built by
`make -C oracle ref`) on a B200.  Test infrastructure -- not product code.

    gpurun -- python tests/golden/gen_reference_golden.py --out gpurun_out/golden
    cp gpurun_out/golden/ref_*.npz tests/golden/          # then commit

Each (case, precision) runs in its own process: the fp64 and fp32 builds share the module name
`PASIfgpu`, and the reference calls std::exit(0) on any CUDA error (src/helpers.cuh:43-51).
Inputs are rebuilt by name from pasif_b200.workloads.golden_case(), the same function the parity
tests use, so a fixture stores outputs + wall times only.
"""
import argparse
import json
import os
import subprocess
import sys
import time

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = os.path.join(ROOT, "oracle", "_ref")


def run_case(name, prec, out_dir):
    sys.path.insert(0, os.path.join(REF, prec))      # PASIfgpu.so of that precision
    sys.path.insert(0, REF)                          # PASIf.py (unmodified copy) + mpi4py shim
    sys.path.insert(0, ROOT)
    import numpy as np
    import PASIf as ref                               # the reference front end
    from pasif_b200 import workloads as wl

    case = wl.golden_case(name)
    exc = [np.array(r, dtype=np.float64) for r in case["excitations"]]
    meta = dict(case=name, prec=prec, E=len(exc), L=len(exc[0]), sample_rate=case["sample_rate"],
                numsteps=case["numsteps"], times={})
    out = {}

    def timed(key, f, *a):
        t = time.time()
        r = f(*a)
        meta["times"][key] = time.time() - t
        return r

    p = ref.PASIf(exc, case["sample_rate"], case["numsteps"])
    timed("setSystems", p.setSystems, *wl.to_frontend_args(case["systems"], ref))
    if case["interp"] is not None:
        table, W = case["interp"]
        p.setInterpolationMatrix([np.array(r) for r in table])
    if name in wl.CALL_ORDER_CASES:                  # repeated solver calls on one driver (quirk A7)
        out = wl.run_call_sequence(p, case, ref)
        meta["gpu"] = subprocess.check_output(["nvidia-smi", "--query-gpu=name,driver_version", "--format=csv,noheader"], text=True).strip()
        meta["reference_commit"] = "5fd9fd1"
        np.savez_compressed(os.path.join(out_dir, f"ref_{name}_{prec}.npz"), meta=json.dumps(meta), **out)
        print("WROTE", name, prec, sorted(out))
        return
    out["amplitudes"] = np.array(timed("getAmplitudes", p.getAmplitudes), dtype=np.float64)
    out["amplitudes_2nd"] = np.array(timed("getAmplitudes_2nd", p.getAmplitudes), dtype=np.float64)
    if case["traj_save_steps"]:
        out["trajectory"] = np.array(timed("getTrajectory", p.getTrajectory, case["traj_save_steps"]), dtype=np.float64)
    if case["jacobians"] is not None:
        timed("setJacobian", p.setJacobian, *wl.to_frontend_args(case["jacobians"], ref, with_psi=True))
        out["gradient"] = np.array(timed("getGradient", p.getGradient), dtype=np.float64)
    try:
        meta["gpu"] = subprocess.check_output(["nvidia-smi", "--query-gpu=name,driver_version", "--format=csv,noheader"],
                                              text=True).strip()
    except Exception as e:  # noqa
        meta["gpu"] = "unknown"
    meta["reference_commit"] = "5fd9fd1"
    np.savez_compressed(os.path.join(out_dir, f"ref_{name}_{prec}.npz"), meta=json.dumps(meta), **out)
    print("WROTE", name, prec, json.dumps(meta["times"]))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default=os.path.join(ROOT, "gpurun_out", "golden"))
    ap.add_argument("--case")
    ap.add_argument("--prec", default="f64")
    ap.add_argument("--cases", default="")
    a = ap.parse_args()
    os.makedirs(a.out, exist_ok=True)
    if a.case:
        run_case(a.case, a.prec, a.out)
        return
    sys.path.insert(0, ROOT)
    from pasif_b200.workloads import GOLDEN_CASES, CALL_ORDER_CASES
    cases = a.cases.split(",") if a.cases else GOLDEN_CASES + CALL_ORDER_CASES
    rc = 0
    for prec in ("f64", "f32"):
        for name in cases:
            log = os.path.join(a.out, f"ref_{name}_{prec}.log")
            with open(log, "w") as fh:
                try:
                    r = subprocess.run([sys.executable, os.path.abspath(__file__), "--out", a.out, "--case", name,
                                        "--prec", prec], stdout=fh, stderr=subprocess.STDOUT, timeout=900)
                    code = r.returncode
                except subprocess.TimeoutExpired:
                    code = -9
            ok = os.path.exists(os.path.join(a.out, f"ref_{name}_{prec}.npz"))
            print(f"{name:16s} {prec} rc={code} fixture={'yes' if ok else 'NO'}", flush=True)
            if not ok:
                rc = 1
                print(open(log).read()[-1500:])
    sys.exit(rc)


if __name__ == "__main__":
    main()
