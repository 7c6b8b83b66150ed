#!/usr/bin/env python3
"""bench.py -- fwd+adjoint gradient evals/s and RK4 system-steps/s of the hot path on N B200s (one JSON line).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload c3|c2|c5] [--impl reference]

A "step" is one pass of the hot path over one batch: `getGradient` (checkpointed forward sweep + per-chunk recompute
+ adjoint sweep) for every (system x excitation) pair of the workload.  K forward-only passes (`getAmplitudes`) are
timed in the same run and reported under "forward" (RK4 system-steps/s).

  value   gradient evals/s with the excitation set already resident in HBM and the result left on the device,
          timed with CUDA events on the engine's stream (sum of K steps), max over ranks.
  e2e     the same metric through the C ABI with HOST buffers: every step uploads the excitation set
          (pasif_load_excitations: H2D inside the timed region) and reads the [E][N_adj] result back to the host.
  N > 1   weak scaling: every rank owns the full per-GPU workload (its own block of systems); the only collective is
          one NCCL all-reduce(SUM) of the excitation-summed adjoint state per step, inside the timed region.

Workloads (BASELINE.json configs):
  c3 (default)  4096 jittered 8-state Gamma+Lambda systems x 64 excitations x 100 000 RK4 steps, fp32, fwd + adjoint
                ("... on 1 B200", the single-GPU throughput configuration)
  c2            1 system x 1024 excitations x 78 001 steps, fp64 (1024 trajectories: latency-, not throughput-bound)
  c5            16384 systems x 64 excitations x 78 001 steps over 8 GPUs, fp64 -> 2048 systems per GPU
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

WORKLOADS = {
    "c3": dict(n_systems=4096, E=64, L=100000, dtype="f32", jitter_seed=1, exc_seed=2,
               name="C3: 4096 jittered 8-state Gamma+Lambda systems x 64 excitations x 100000 RK4 steps, fp32, forward + adjoint"),
    "c2": dict(n_systems=1, E=1024, L=78001, dtype="f64", jitter_seed=None, exc_seed=0,
               name="C2: 8-state Gamma+Lambda system x 1024 excitations x 78001 RK4 steps, fp64, forward + adjoint"),
    "c5": dict(n_systems=2048, E=64, L=78001, dtype="f64", jitter_seed=4, exc_seed=2,
               name="C5 shard: 2048 jittered 8-state systems x 64 excitations x 78001 RK4 steps per GPU, fp64, forward + adjoint"),
    # config 4 (large-DOF shared-B batch): ONE 256-DOF chain (514 states, 518-state adjoint) x 16384 excitation replicas.
    # The rate is step-count independent; the step counts keep a bench step at a few seconds.
    "c4s": dict(kind="chain", dense=False, ndof=256, n_systems=1, E=16384, L=2000, dtype="f64", jitter_seed=None, exc_seed=5,
                name="C4 sparse: 256-DOF oscillator chain (tridiagonal K, on-site Gamma/Lambda; 514 states) x 16384 excitations x 2000 RK4 steps, forward + adjoint"),
    "c4d": dict(kind="chain", dense=True, ndof=256, n_systems=1, E=16384, L=128, dtype="f64", jitter_seed=None, exc_seed=5,
                name="C4 dense: 256-DOF system with dense shared K, B (514 states) x 16384 excitations x 128 RK4 steps, forward + adjoint"),
}

# dram__bytes_read.sum + dram__bytes_write.sum per launch of the hot kernels, from one `ncu --set full` capture of the same
# command at the workload's full size (profiles/r01_ncu_c3_full_config_kernels.csv: chunk of 317 steps x 262144 trajectories;
# algorithmic bytes of that launch: 317 x 262144 x 4 needed variables x 4 B = 1.330e9)
NCU_TRAFFIC = {("c3", "f32", "reference"): {"gfwd": 1.270203e9 + 0.008973e9, "bwd": 1.346701e9 + 0.013820e9,
                                            "source": "profiles/r01_ncu_c3_full_config_kernels.csv"}}

# algorithmic flops per trajectory-step of the 8-state system / 12-state Jacobian:
#   D = 2 nnzK + 3 nnzG + 4 nnzL + 5 nnzP + 2 nnzF ; forward step 4 D + 13 N ; adjoint step 4 D_adj + 13 N_adj + 2 N
F_FWD, F_BWD = 240.0, 380.0
N_FWD, N_ADJ, N_NEEDED = 8, 12, 4


def system_work(s, j):
    """(F_fwd, F_bwd, N_fwd, N_adj) of one system / Jacobian pair from their non-zero counts."""
    def nnzk(d):
        rows = list(map(int, d["B"][1])) + list(map(int, d["K"][1]))
        cols = list(map(int, d["B"][2])) + list(map(int, d["K"][2]))
        return len(set(zip(rows, cols)))

    def D(d, force):
        return (2 * nnzk(d) + 3 * len(d["Gamma"][1]) + 4 * len(d["Lambda"][1]) + 5 * len(d.get("Psi", ([], [], []))[1])
                + (2 * int(np.count_nonzero(d["F"])) if force else 0))
    return (4.0 * D(s, True) + 13 * s["n"], 4.0 * D(j, False) + 13 * j["n"] + 2 * s["n"], s["n"], j["n"])


def build_systems(w, rank):
    from pasif_b200 import workloads as wl
    if w.get("kind") == "chain":
        s = wl.chain_system(w["ndof"], dense=w["dense"])
        return [s], [wl.chain_jacobian(s, 4)]
    if w["jitter_seed"] is None:
        return [wl.fit8_system() for _ in range(w["n_systems"])], [wl.fit8_jacobian() for _ in range(w["n_systems"])]
    rng = np.random.default_rng(w["jitter_seed"] + 1000 * rank)
    systems = [wl.jitter(wl.fit8_system(), rng) for _ in range(w["n_systems"])]
    jacs = [wl.fit8_jacobian() for _ in range(w["n_systems"])]
    return systems, jacs


def schedule(n, dtype):
    """chunk schedule of the literal reference mode (reference src/__GpuDriver.cu:506-508)"""
    t = np.float32 if dtype == "f32" else np.float64
    nsp = int(np.sqrt(float(n)))
    c = int(np.ceil(t(n) / t(nsp)))
    return nsp, c


class ClockSampler(threading.Thread):
    """nvidia-smi clocks + throttle reasons during the timed region (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu):
        super().__init__(daemon=True)
        self.gpu, self.rows, self._stop = gpu, [], threading.Event()

    def run(self):
        try:
            self.p = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "200",
                                       "-i", str(self.gpu)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            for line in self.p.stdout:
                self.rows.append([x.strip() for x in line.split(",")])
                if self._stop.is_set():
                    break
        except Exception:
            pass

    def stop(self):
        self._stop.set()
        try:
            self.p.terminate()
        except Exception:
            pass
        sm = [float(r[1]) for r in self.rows if len(r) >= 9 and r[1].replace(".", "").isdigit()]
        mx = [float(r[2]) for r in self.rows if len(r) >= 9 and r[2].replace(".", "").isdigit()]
        reasons = set()
        for r in self.rows:
            if len(r) >= 9:
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def cpu_baseline_sample(w, threads, seconds_target=15.0):
    """The CPU oracle (C port of the reference's algorithm; the reference has no CPU path) on a bounded sample of the
    same workload: a few systems x all excitations x a reduced step count.  Returns (gradient evals/s equivalent at the
    workload's step count, forward system-steps/s, description)."""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    from pasif_oracle_c import CRecordingDriver
    from pasif_b200 import frontend, workloads as wl

    class mod:
        coo_tensor = frontend.coo_tensor

    class D(CRecordingDriver):
        dtype = np.float32 if w["dtype"] == "f32" else np.float64
    if w.get("kind") == "chain":
        n_sys, E = 1, min(w["E"], 4 * threads)
        steps = 40 if w["dense"] else 400
        systems, jacs = build_systems(w, 0)
        what = w["name"].split(":")[0]
    else:
        n_sys = min(w["n_systems"], 16)
        E = min(w["E"], 64)
        steps = w["L"]                         # full length: ~10 s of work on 16 host cores
        rng = np.random.default_rng(1)
        systems = [wl.jitter(wl.fit8_system(), rng) for _ in range(n_sys)]
        jacs = [wl.fit8_jacobian() for _ in range(n_sys)]
        what = f"{n_sys} systems"
    exc = wl.random_sine_excitations(E, steps, 16000, seed=w["exc_seed"])
    p = frontend.make_frontend(D)([np.array(r) for r in exc], 16000, 0)
    p._oracle.threads = threads
    p.setSystems(*wl.to_frontend_args(systems, mod))
    p.setJacobian(*wl.to_frontend_args(jacs, mod, with_psi=True))
    t0 = time.time()
    p.getAmplitudes()
    t_f = time.time() - t0
    t0 = time.time()
    p.getGradient()
    t_g = time.time() - t0
    pairs = n_sys * E
    grad_traj_steps_per_s = pairs * steps / t_g
    return (grad_traj_steps_per_s / w["L"], pairs * steps / t_f,
            f"{what} x {E} excitations x {steps} steps ({w['dtype']}), C oracle, OpenMP {threads} threads; "
            f"evals/s scaled to {w['L']} steps", t_f + t_g)


def reference_gpu_sample(w):
    """The reference's own (GPU) implementation, unmodified, on the same B200: oracle/time_reference_gpu.py in a
    subprocess.  Reported beside the CPU figures; None when no GPU / no compiled reference / not an 8-state workload."""
    so = os.path.join(ROOT, "oracle", "_ref", w["dtype"], "PASIfgpu.so")
    if w.get("kind") == "chain" or not os.path.exists(so):
        return None
    try:
        if subprocess.run(["nvidia-smi", "-L"], capture_output=True, timeout=30).returncode != 0:
            return None
        S, E, steps = min(w["n_systems"], 2), min(w["E"], 64), 20000
        r = subprocess.run([sys.executable, os.path.join(ROOT, "oracle", "time_reference_gpu.py"), "--prec", w["dtype"],
                            "--systems", str(S), "--excitations", str(E), "--steps", str(steps), "--exc-seed", str(w["exc_seed"])],
                           capture_output=True, text=True, timeout=600)
        d = json.loads(r.stdout.strip().splitlines()[-1])
        return {"value": d["gradient_trajectory_steps_per_s"] / w["L"], "unit": "gradient evals/s",
                "forward_system_steps_per_s": d["forward_system_steps_per_s"], "gradient_s": d["gradient_s"], "forward_s": d["forward_s"],
                "sample": f"unmodified reference (sm_100 build of src/*.cu, {w['dtype']}) through its own PASIf.py on this GPU: {S} systems x {E} "
                          f"excitations x {steps} steps (its front end cannot assemble more than 2 Jacobians; its per-step cost is "
                          f"launch-bound, i.e. independent of the batch at this size); evals/s scaled to {w['L']} steps"}
    except Exception as e:  # noqa
        return {"unavailable": repr(e)[:200]}


def run_reference_arm(a, w):
    """--impl reference: the reference ships no CPU implementation of this path, so this arm times
    the oracle port -- a literal C restatement validated against the compiled reference's B200 outputs -- on the host
    cores, each step one bounded sample of the workload."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    vals, fvals = [], []
    sample = ""
    for i in range(a.warmup + a.steps):
        v, fv, sample, _ = cpu_baseline_sample(w, threads)
        if i >= a.warmup:
            vals.append(v)
            fvals.append(fv)
    value = float(np.mean(vals))
    # a step of the native arm is one gradient of the whole batch; the CPU needs pairs / value seconds for it (extrapolated
    # from the bounded sample: the rate does not depend on the batch size)
    pairs = w["n_systems"] * w["E"] * max(a.gpus, 1)
    line = {"metric": "fwd+adjoint gradient evals/s", "value": value, "unit": "gradient evals/s", "n_gpus": a.gpus,
            "steps": a.steps, "warmup": a.warmup, "ms_per_step": pairs / value * 1e3 if value > 0 else None,
            "ms_per_step_note": "whole-batch step extrapolated from the bounded sample",
            "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": w["dtype"], "data": "synthetic", "impl": "reference",
            "config": {"workload": w["name"]},
            "cpu_baseline": {"value": value, "unit": "gradient evals/s", "cores": threads, "kind": "port", "sample": sample},
            "forward": {"metric": "RK4 system-steps/s", "value": float(np.mean(fvals)), "unit": "system-steps/s"},
            "e2e": {"value": value, "unit": "gradient evals/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    rg = reference_gpu_sample(w)
    if rg is not None:
        line["reference_gpu"] = rg
    print(json.dumps(line))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--workload", default="c3", choices=sorted(WORKLOADS))
    ap.add_argument("--impl", default="native", choices=["native", "reference"])
    ap.add_argument("--systems", type=int, default=0, help="override the number of systems per GPU (debug)")
    ap.add_argument("--length", type=int, default=0, help="override the number of RK4 steps (debug)")
    ap.add_argument("--dtype", default="", choices=["", "f32", "f64"], help="override the workload's compute type")
    ap.add_argument("--excitations", type=int, default=0, help="override the number of excitations (debug)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--compat", default="reference", choices=["reference", "clean"],
                    help="adjoint schedule: the reference's literal one (default, parity with its outputs) or the exact "
                         "discrete adjoint, whose chunk trajectories are kept from the forward sweep as far as HBM allows")
    ap.add_argument("--tail-budget-gb", default="auto", help="clean schedule: device memory for kept chunk trajectories")
    ap.add_argument("--option", action="append", default=[], help="engine option key=value (debug), repeatable")
    ap.add_argument("--stage-parts", default="", help="dense-operator path: number of batch parts on their own streams (debug)")
    ap.add_argument("--lane-regs", default="", help="register estimate up to which a lane pattern is accepted (debug)")
    ap.add_argument("--lanes", default="on", help="lane-block split of large components: on | off | widest group (8..128)")
    a = ap.parse_args()
    w = dict(WORKLOADS[a.workload])
    if a.systems:
        w["n_systems"] = a.systems
    if a.length:
        w["L"] = a.length
    if a.excitations:
        w["E"] = a.excitations
    if a.systems or a.length or a.excitations:
        w["name"] += f" [override: {w['n_systems']} systems, {w['E']} excitations, {w['L']} steps]"
    if a.dtype:
        w["dtype"] = a.dtype
        w["name"] += f" [{a.dtype}]"
    if a.impl == "reference":
        return run_reference_arm(a, w)

    # stdout carries exactly ONE line (the JSON): anything a library prints there while the run is set up (NCCL's version
    # banner under NCCL_DEBUG=VERSION, for one) is sent to stderr instead
    sys.stdout.flush()
    real_stdout = os.dup(1)
    os.dup2(2, 1)

    import torch
    from pasif_b200 import capi, frontend, workloads as wl
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    dist = None
    if world > 1:
        import torch.distributed as dist
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        torch.cuda.set_device(local_rank)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    torch.cuda.set_device(local_rank)

    class mod:
        coo_tensor = frontend.coo_tensor

    S, E, L = w["n_systems"], w["E"], w["L"]
    systems, jacs = build_systems(w, rank)
    global F_FWD, F_BWD, N_FWD, N_ADJ, N_NEEDED
    F_FWD, F_BWD, N_FWD, N_ADJ = system_work(systems[0], jacs[0])
    large = w.get("kind") == "chain"
    # forward variables written once (recompute) and read once (adjoint): the needed ones (register path) or all rows
    # padded to 32 (tile path)
    N_NEEDED = (N_FWD + 31) // 32 * 32 if large else 4
    exc = np.ascontiguousarray(wl.random_sine_excitations(E, L, 16000, seed=w["exc_seed"]))
    P = frontend.make_frontend(capi.engine_class(w["dtype"]))

    class Local(P):
        pass
    import pasif_b200.frontend as fe
    saved = fe._local_rank
    fe._local_rank = lambda: local_rank
    p = Local(exc, 16000, 0)
    fe._local_rank = saved
    p._setOption("compat", a.compat)
    p._setOption("tail_budget_gb", a.tail_budget_gb)
    p._setOption("lanes", a.lanes)
    if a.lane_regs:
        p._setOption("lane_regs", a.lane_regs)
    if a.stage_parts:
        p._setOption("stage_parts", a.stage_parts)
    for kv in a.option:
        k, v = kv.split("=", 1)
        p._setOption(k, v)
    t_setup = time.time()
    p.setSystems(*wl.to_frontend_args(systems, mod))
    p.setJacobian(*wl.to_frontend_args(jacs, mod, with_psi=True))
    t_setup = time.time() - t_setup
    tdt = torch.float32 if w["dtype"] == "f32" else torch.float64
    red = torch.zeros(world * S * N_ADJ, dtype=tdt, device=f"cuda:{local_rank}")
    my = red[rank * S * N_ADJ:(rank + 1) * S * N_ADJ]

    def barrier():
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    def grad_step_resident():
        if dist is None:
            p._getGradient(0, fetch=False)
            return p._query("last_device_ms")
        p._getGradientReducedDevice(my.data_ptr(), 0)
        ms = p._query("last_device_ms")
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ev0.record()
        dist.all_reduce(red)
        ev1.record()
        torch.cuda.synchronize()
        return ms + ev0.elapsed_time(ev1)

    pairs = S * E
    # ---------------- warm-up (>= 3), then K timed gradient steps with resident inputs
    for _ in range(max(a.warmup, 3)):
        grad_step_resident()
        p._getAmplitudes(fetch=False)
    sampler = ClockSampler(local_rank)
    sampler.start()
    barrier()
    wall0 = time.time()
    dev_ms = 0.0
    launches = 0
    for _ in range(a.steps):
        dev_ms += grad_step_resident()
        launches += int(p._query("last_launches"))
    barrier()
    wall_grad = time.time() - wall0
    # ---------------- per-kernel device times: K more steps with profile=on (a CUDA-event pair around every hot-kernel
    # launch; the recompute and adjoint kernels are serialised for this pass, they overlap on two streams above)
    k_ms = {"gfwd": 0.0, "bwd": 0.0}
    k_launch = {"gfwd": 0, "bwd": 0}
    u_ms = {"upd_fwd": 0.0, "upd_adj": 0.0}          # dense-operator path only: its fused update kernels
    u_launch = {"upd_fwd": 0, "upd_adj": 0}
    p._setOption("profile", "on")
    serial_ms = 0.0
    for _ in range(a.steps):
        p._getGradient(0, fetch=False)
        serial_ms += p._query("last_device_ms")
        for k in k_ms:
            k_ms[k] += p._query("ms_" + k)
            k_launch[k] += int(p._query("launches_" + k))
        for k in u_ms:
            u_ms[k] += p._query("ms_" + k)
            u_launch[k] += int(p._query("launches_" + k))
    p._setOption("profile", "off")
    barrier()
    # ---------------- K forward-only passes (RK4 system-steps/s)
    fwd_ms = 0.0
    for _ in range(a.steps):
        p._getAmplitudes(fetch=False)
        fwd_ms += p._query("last_device_ms")
        launches += int(p._query("last_launches"))
    barrier()
    clocks = sampler.stop()
    # ---------------- e2e: host buffers in, host result out, every step
    e2e_t = 0.0
    e2e_parts = {"load_excitations_ms": 0.0, "h2d_staging_ms": 0.0, "gradient_call_ms": 0.0, "device_ms": 0.0, "d2h_staging_ms": 0.0}
    for i in range(1 + a.steps):
        barrier()
        t0 = time.time()
        p._loadExcitationsSet(exc, 16000)
        t1 = time.time()
        g = p._getGradient(0)
        t2 = time.time()
        if dist is not None:
            dist.all_reduce(red)
            torch.cuda.synchronize()
        if i > 0:
            e2e_t += time.time() - t0
            e2e_parts["load_excitations_ms"] += (t1 - t0) * 1e3 / a.steps
            e2e_parts["gradient_call_ms"] += (t2 - t1) * 1e3 / a.steps
            e2e_parts["device_ms"] += p._query("last_device_ms") / a.steps
            e2e_parts["h2d_staging_ms"] += p._query("last_h2d_ms") / a.steps
            e2e_parts["d2h_staging_ms"] += p._query("last_d2h_ms") / a.steps
    tmax = torch.tensor([dev_ms, fwd_ms, e2e_t, wall_grad], dtype=torch.float64, device=f"cuda:{local_rank}")
    if dist is not None:
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
    dev_ms, fwd_ms, e2e_t, wall_grad = [float(x) for x in tmax.tolist()]
    if rank != 0:
        if dist is not None:
            dist.destroy_process_group()
        return

    tot_pairs = pairs * world
    value = tot_pairs * a.steps / (dev_ms * 1e-3)
    fwd_value = tot_pairs * L * a.steps / (fwd_ms * 1e-3)
    e2e_value = tot_pairs * a.steps / e2e_t
    nsp, c = schedule(L, w["dtype"])
    stored_chunks = 0
    if a.compat == "reference":
        # literal reference schedule: forward sweep n steps + recompute (n - c) steps; adjoint (n - c) steps (chunk 0 skipped)
        steps_gfwd, steps_bwd = 2 * L - c, L - c
    else:
        # exact schedule: sweep n steps, recompute only the chunks whose trajectory did not fit in HBM, adjoint n steps
        c = int(np.ceil(np.sqrt(L)))
        nsp = (L + c - 1) // c
        stored_chunks = int(p._query("stored_chunks"))
        steps_gfwd, steps_bwd = L + min(L, (nsp - stored_chunks) * c), L
    flops_gfwd = pairs * steps_gfwd * F_FWD
    flops_bwd = pairs * steps_bwd * F_BWD
    fma = {"f64": capi.fma_peak(local_rank, "f64"), "f32": capi.fma_peak(local_rank, "f32")}
    peak = fma[w["dtype"]]
    per_kernel = {}
    for k, fl in (("gfwd", flops_gfwd), ("bwd", flops_bwd)):
        t = k_ms[k] / a.steps * 1e-3
        per_kernel[k] = {"tflops": fl / t / 1e12 if t > 0 else None, "ms_per_step": k_ms[k] / a.steps,
                         "launches_per_step": k_launch[k] // a.steps,
                         "avg_launch_ms": k_ms[k] / max(k_launch[k], 1)}
    dom = max(per_kernel, key=lambda k: per_kernel[k]["ms_per_step"])
    sz = 4 if w["dtype"] == "f32" else 8
    peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))) if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else {}
    stage = large and p._query("large_stage_classes") > 0
    lane = p._query("lane_classes") > 0
    if lane:
        # lane-block split: every lane writes / reads its share of the needed forward variables
        N_NEEDED = int(p._query("lane_nv") * p._query("lane_G"))
    tensor = None
    if stage:
        # dense-operator path: the profiled kinds time the per-stage GEMM launches only (one GEMM per RK4 stage).
        # 3xTF32 issues three tensor-core products per useful one.
        pad = lambda n, m: (n + m - 1) // m * m
        mult = 3 if w["dtype"] == "f32" else 1
        ppad = pad(pairs, 128)
        # the GEMM covers the band of dense rows of K' (x' = v rows and probes stay row programs): rows x padded columns
        gemm_flops = {"gfwd": 2.0 * p._query("stage_band_rows_fwd") * p._query("stage_band_cols_fwd") * ppad * mult,
                      "bwd": 2.0 * p._query("stage_band_rows_adj") * p._query("stage_band_cols_adj") * ppad * mult}
        if w["dtype"] == "f32":
            tpeak = peaks.get("bf16_tflops", 1590.0) / 2.0
            tsrc = "tf32 dense = half the bf16 rate: MEASURED_PEAKS.json bf16_tflops / 2" if peaks else "fallback 1590 / 2"
            tname = "sgemm_tn_3xtf32_tcgen05_kernel (tcgen05.mma kind::tf32, TMEM accumulator, TMA operands)"
        else:
            tpeak = capi.fma_peak(local_rank, "dmma")
            tsrc = "register-resident mma.sync.m8n8k4.f64 microbenchmark, same run (tcgen05 has no fp64 kind)"
            tname = "dgemm_tn_dmma_kernel (mma.sync.m8n8k4.f64, cp.async pipeline)"
        ach = gemm_flops[dom] / (per_kernel[dom]["avg_launch_ms"] * 1e-3) / 1e12
        tensor = {"bound": "tensor", "kernel": tname, "achieved": ach, "peak": tpeak, "unit": "TFLOP/s", "frac": ach / tpeak,
                  "traffic": None, "peak_source": tsrc, "tensor_products_per_useful_product": mult,
                  "useful_tflops": ach / mult, "gemm_us_per_launch": per_kernel[dom]["avg_launch_ms"] * 1e3,
                  "gemm_rows": {"fwd": p._query("stage_band_rows_fwd"), "adj": p._query("stage_band_rows_adj")},
                  "gemm_cols": {"fwd": p._query("stage_band_cols_fwd"), "adj": p._query("stage_band_cols_adj")},
                  "gemm_share_of_step": (k_ms["gfwd"] + k_ms["bwd"]) / max(serial_ms, 1e-9),
                  "whole_step": {"achieved": (flops_gfwd + flops_bwd) / (dev_ms / a.steps * 1e-3) / 1e12,
                                 "frac": (flops_gfwd + flops_bwd) / (dev_ms / a.steps * 1e-3) / 1e12 / (tpeak / mult)},
                  "per_kernel": per_kernel, "serialised_ms_per_step": serial_ms / a.steps,
                  "note": "per-GEMM time from a serialised pass (profile=on); value/whole_step from the overlapped two-part schedule; "
                          "whole_step.frac is against peak / tensor_products_per_useful_product"}
    # chunk-trajectory traffic: needed forward variables written once (recompute) and read once (adjoint)
    hbm_bytes = pairs * steps_bwd * 2 * N_NEEDED * sz + pairs * nsp * (N_NEEDED if large else N_FWD) * sz * 2
    hbm_peak = peaks.get("hbm_gbs", 6650.0)
    hbm_note = "algorithmic chunk-trajectory + checkpoint bytes of the whole gradient step over its device time"
    stage_hbm = None
    if stage:
        # dense-operator path: the state lives in HBM between the stages.  Algorithmic values moved per trajectory and
        # RK4 stage (averaged over the four stages): the GEMM reads the stage vector (band columns) and writes Y (band
        # rows); the update reads Y (band rows), Q, S (stages 2-4), the stage vector (row-program operands; just read by
        # the GEMM, so partly L2 hits) and, in the adjoint, the forward set point (twice on the half stages); it writes
        # S (stages 1-3), the next stage vector and, in stage 4, Q (+ the chunk-trajectory slot while recomputing).
        q = lambda k: p._query("stage_band_" + k)
        nrf, nra = q("nr_fwd"), q("nr_adj")
        # stage_direct (default): stage 1 evaluates at the state array (no separate stage vector to read) and stage 4 writes
        # the state array only: Q 1 + S .75 + stage vector .75 read; S .75 + stage vector .75 + Q .25 written
        per_nr = (1 + .75 + .75) + (.75 + .75 + .25) if p._query("stage_direct") else (1 + .75 + 1) + (.75 + 1 + .25)
        upd_f = q("rows_fwd") + per_nr * nrf
        upd_a = q("rows_adj") + per_nr * nra + 1.5 * nrf
        gemm_f, gemm_a = q("cols_fwd") + q("rows_fwd"), q("cols_adj") + q("rows_adj")
        stages_f, stages_a = 4 * steps_gfwd, 4 * steps_bwd
        hbm_bytes = pairs * sz * (stages_f * (upd_f + gemm_f) + (steps_gfwd - L) * nrf + stages_a * (upd_a + gemm_a))
        hbm_note = ("dense-operator path: algorithmic stage-vector bytes (GEMM operand / result + the update kernel's Y, Q, S, "
                    "stage vector, set point) of the whole gradient step over its device time")
        ppad_u = (pairs + 31) // 32 * 32
        stage_hbm = {}
        for k, vals in (("upd_fwd", upd_f), ("upd_adj", upd_a)):
            if u_launch[k]:
                us = u_ms[k] / u_launch[k] * 1e3
                gbs = pairs * vals * sz / (us * 1e-6) / 1e9
                stage_hbm[k] = {"kernel": "stage_update_kernel", "us_per_launch": us, "launches_per_step": u_launch[k] // a.steps,
                                "ms_per_step": u_ms[k] / a.steps, "algorithmic_bytes_per_launch": pairs * vals * sz,
                                "achieved_gbs": gbs, "frac_of_hbm_peak": gbs / hbm_peak}
    line = {
        "metric": "fwd+adjoint gradient evals/s", "value": value, "unit": "gradient evals/s", "n_gpus": world,
        "steps": a.steps, "warmup": max(a.warmup, 3), "ms_per_step": dev_ms / a.steps, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": w["dtype"], "data": "synthetic",
        "config": {"workload": w["name"], "systems_per_gpu": S, "excitations": E, "rk4_steps": L, "compat": a.compat,
                   "schedule": {"chunks": nsp, "chunk_steps": c, "chunks_kept_in_hbm": stored_chunks,
                                "forward_steps_per_eval": steps_gfwd, "adjoint_steps_per_eval": steps_bwd},
                   "l2": "working set (checkpoints + chunk trajectory, > 1 GB) exceeds the 126 MB L2; no flush needed",
                   "timing": "CUDA events on the engine stream, summed over the timed steps, max over ranks",
                   "setup_s": t_setup},
        "forward": {"metric": "RK4 system-steps/s", "value": fwd_value, "unit": "system-steps/s",
                    "ms_per_step": fwd_ms / a.steps,
                    "roofline": {"bound": "fma", "achieved": pairs * L * F_FWD / (fwd_ms / a.steps * 1e-3) / 1e12,
                                 "peak": peak, "unit": "TFLOP/s",
                                 "frac": pairs * L * F_FWD / (fwd_ms / a.steps * 1e-3) / 1e12 / peak}},
        "roofline": {"bound": "fma", "kernel": ({"gfwd": "large_forward_kernel", "bwd": "large_backward_kernel"} if (large and not lane) else {"gfwd": "pasif_gfwd_*", "bwd": "pasif_bwd_*"})[dom],
                     "achieved": per_kernel[dom]["tflops"], "peak": peak, "unit": "TFLOP/s",
                     "frac": per_kernel[dom]["tflops"] / peak if per_kernel[dom]["tflops"] else None, "traffic": None,
                     "peak_source": "pasif_fma_peak register-resident FMA microbenchmark, same run (MEASURED_PEAKS.json has no FMA-pipe figure)",
                     "whole_step": {"achieved": (flops_gfwd + flops_bwd) / (dev_ms / a.steps * 1e-3) / 1e12,
                                    "frac": (flops_gfwd + flops_bwd) / (dev_ms / a.steps * 1e-3) / 1e12 / peak},
                     "per_kernel": per_kernel, "fma_peak_tflops": fma,
                     "serialised_ms_per_step": serial_ms / a.steps,
                     "note": "per-kernel figures from a serialised pass (profile=on); value/whole_step from the two-stream overlapped schedule"},
        "roofline_hbm": {"bound": "hbm", "achieved": hbm_bytes / (dev_ms / a.steps * 1e-3) / 1e9, "peak": hbm_peak,
                         "unit": "GB/s", "frac": hbm_bytes / (dev_ms / a.steps * 1e-3) / 1e9 / hbm_peak,
                         "note": hbm_note,
                         "peak_source": "MEASURED_PEAKS.json" if peaks else "fallback"},
        "e2e": {"value": e2e_value, "unit": "gradient evals/s", "h2d_bytes_per_step": int(exc.nbytes),
                "d2h_bytes_per_step": int(pairs * N_ADJ * sz), "host_buffers": "pageable numpy arrays through the C ABI, staged through two pinned slots inside the engine",
                "ms_per_step": e2e_t / a.steps * 1e3, "parts": e2e_parts},
        "gpu_launches": launches, "clocks": clocks, "wall_s_grad_steps": wall_grad,
    }
    tr = NCU_TRAFFIC.get((a.workload, w["dtype"], a.compat))
    if tr and not (a.systems or a.length or a.excitations) and not large:
        line["roofline"]["traffic"] = tr[dom]
        line["roofline"]["traffic_source"] = tr["source"] + " (ncu --set full, same command; bytes per launch of the dominant kernel)"
        line["roofline"]["algorithmic_bytes_per_launch"] = pairs * c * N_NEEDED * sz
    if lane:
        line["config"]["lane_split"] = {k: p._query("lane_" + k) for k in ("G", "R", "RA", "nv", "nc_fwd", "nc_adj", "halo_fwd", "halo_adj", "halo_p", "efficiency")}
        line["config"]["path"] = "lane-block register kernels (lane_split.hpp): G threads per trajectory, state in registers, halo exchange per stage"
    elif large:
        line["config"]["path"] = "dense-operator stage path (tensor-core GEMM per stage)" if stage else "shared-memory tile kernels"
    if stage_hbm:
        line["roofline_hbm"]["update_kernels"] = stage_hbm
    if tensor is not None:
        line["roofline_fma"] = line["roofline"]
        line["roofline"] = tensor
        line["forward"]["roofline"] = {"bound": "tensor", "achieved": line["forward"]["roofline"]["achieved"], "unit": "TFLOP/s",
                                       "peak": tensor["peak"] / tensor["tensor_products_per_useful_product"],
                                       "frac": line["forward"]["roofline"]["achieved"] / (tensor["peak"] / tensor["tensor_products_per_useful_product"])}
    if not a.no_cpu_baseline and world == 1:
        threads = os.cpu_count() or 1
        v, fv, sample, _ = cpu_baseline_sample(w, threads)
        line["cpu_baseline"] = {"value": v, "unit": "gradient evals/s", "cores": threads, "kind": "port", "sample": sample,
                                "forward_system_steps_per_s": fv}
    sys.stdout.flush()
    os.write(real_stdout, (json.dumps(line) + "\n").encode())
    if dist is not None:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
