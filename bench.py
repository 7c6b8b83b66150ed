#!/usr/bin/env python3
"""bench.py -- fwd+adjoint gradient evals/s and RK4 system-steps/s of the hot path on N B200s (one JSON line).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload c3|c1|c2|c4s|c4d|c4r|c5|c5s] [--impl reference]

A "step" is one pass of the hot path over one batch: `getGradient` (checkpointed forward sweep + per-chunk recompute
+ adjoint sweep) for every (system x excitation) pair of the workload.  K forward-only passes (`getAmplitudes`) are
timed in the same run and reported under "forward" (RK4 system-steps/s).

  value      gradient evals/s with the excitation set already resident in HBM and the result left on the device,
             timed with CUDA events on the engine's stream (sum of K steps), max over ranks.
  e2e        the same metric through the C ABI with HOST buffers: every step uploads the excitation set
             (pasif_load_excitations: H2D inside the timed region) and reads the [E][N_adj] result back to the host
             (N > 1: plus the all-reduce of that step's excitation-summed gradient and its read-back).
  N > 1      the headline is weak scaling (every rank owns the full per-GPU workload, its own block of systems) through
             pasif_b200.sharded.ShardedPASIf; the only collective is one NCCL all-reduce(SUM) of the excitation-summed
             adjoint state per step, inside the timed region.
  c5_strong  BASELINE config 5 at its full size (16384 systems x 64 excitations x 78001 steps, fp64), a FIXED total
             sharded over the N ranks (strong scaling), with per-rank device times, the all-reduce time and -- outside
             the timed region -- a check of the all-reduced vector against a single-GPU recompute of a slice.
  configs    (default run, N = 1) one entry per remaining BASELINE config: C1, C2, C4 sparse, C4 dense (fp32 + fp64,
             two step counts), C4 parameter replicas; a few timed steps each.

Workloads (BASELINE.json configs):
  c3 (default)  4096 jittered 8-state Gamma+Lambda systems x 64 excitations x 100 000 RK4 steps, fp32, fwd + adjoint
                ("... on 1 B200", the single-GPU throughput configuration)
  c1            1-DOF Duffing oscillator x 8 sine excitations x 78 001 steps, fp64, forward amplitudes
  c2            1 system x 1024 excitations x 78 001 steps, fp64 (1024 trajectories: latency-, not throughput-bound)
  c4s / c4d     256-DOF chain (514 states) x 16384 excitations, sparse / dense shared operator
  c4r           16 parameter replicas (shared dense K', jittered Gamma / Lambda / F) x 1024 excitations
  c5 / c5s      2048 systems per GPU (weak) / 16384 systems in total (strong) x 64 excitations x 78 001 steps, fp64
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
    "c1": dict(kind="duffing", n_systems=1, E=8, L=78001, dtype="f64", jitter_seed=None, exc_seed=None, forward_only=True,
               name="C1: 1-DOF Duffing oscillator (cubic Lambda, 4 states) x 8 sine excitations x 78001 RK4 steps, fp64, forward amplitudes"),
    "c3": dict(n_systems=4096, E=64, L=100000, dtype="f32", jitter_seed=1, exc_seed=2,
               name="C3: 4096 jittered 8-state Gamma+Lambda systems x 64 excitations x 100000 RK4 steps, fp32, forward + adjoint"),
    "c2": dict(n_systems=1, E=1024, L=78001, dtype="f64", jitter_seed=None, exc_seed=0,
               name="C2: 8-state Gamma+Lambda system x 1024 excitations x 78001 RK4 steps, fp64, forward + adjoint"),
    "c5": dict(n_systems=2048, E=64, L=78001, dtype="f64", jitter_seed=4, exc_seed=2,
               name="C5 shard: 2048 jittered 8-state systems x 64 excitations x 78001 RK4 steps per GPU, fp64, forward + adjoint"),
    "c5s": dict(n_systems=16384, E=64, L=78001, dtype="f64", jitter_seed=4, exc_seed=2, strong=True,
                name="C5: 16384 jittered 8-state systems x 64 excitations (1048576 pairs) x 78001 RK4 steps, fp64, forward + adjoint, "
                     "fixed total sharded over the ranks"),
    # config 4 (large-DOF shared-B batch): ONE 256-DOF chain (514 states, 518-state adjoint) x 16384 excitation replicas
    "c4s": dict(kind="chain", dense=False, ndof=256, n_systems=1, E=16384, L=10000, dtype="f64", jitter_seed=None, exc_seed=5,
                name="C4 sparse: 256-DOF oscillator chain (tridiagonal K, on-site Gamma/Lambda; 514 states) x 16384 excitations x 10000 RK4 steps, forward + adjoint"),
    # the dense operator costs 8-20 s per gradient at 10000 steps: it is timed at two shorter lengths (the rate per
    # trajectory-step does not depend on the length; both are in the record)
    "c4d": dict(kind="chain", dense=True, ndof=256, n_systems=1, E=16384, L=128, dtype="f64", jitter_seed=None, exc_seed=5,
                name="C4 dense: 256-DOF system with dense shared K, B (514 states) x 16384 excitations x 128 RK4 steps, forward + adjoint"),
    "c4r": dict(kind="chain", dense=True, ndof=256, n_systems=16, E=1024, L=128, dtype="f32", jitter_seed=None, exc_seed=5, replica_seed=11,
                name="C4 parameter replicas: 16 instances sharing one dense K' (514 states) with jittered Gamma / Lambda / F x 1024 excitations x 128 RK4 steps, forward + adjoint"),
}

# dram__bytes_read.sum + dram__bytes_write.sum per launch of the hot kernels, from one `ncu --set full` capture of the same
# command at the workload's full size (chunk of 317 steps x 262144 trajectories; algorithmic bytes of that launch:
# 317 x 262144 x 4 needed variables x 4 B = 1.330e9)
NCU_TRAFFIC = {("c3", "f32", "reference"): {"gfwd": 1.270203e9 + 0.008973e9, "bwd": 1.346701e9 + 0.013820e9,
                                            "source": "profiles/r01_ncu_c3_full_config_kernels.csv"}}


def system_work(s, j):
    """(F_fwd, F_bwd, N_fwd, N_adj): algorithmic flops per trajectory-step from the non-zero counts:
    D = 2 nnzK + 3 nnzG + 4 nnzL + 5 nnzP + 2 nnzF ; forward step 4 D + 13 N ; adjoint step 4 D_adj + 13 N_adj + 2 N"""
    def nnzk(d):
        rows = list(map(int, d["B"][1])) + list(map(int, d["K"][1]))
        cols = list(map(int, d["B"][2])) + list(map(int, d["K"][2]))
        return len(set(zip(rows, cols)))

    def D(d, force):
        return (2 * nnzk(d) + 3 * len(d["Gamma"][1]) + 4 * len(d["Lambda"][1]) + 5 * len(d.get("Psi", ([], [], []))[1])
                + (2 * int(np.count_nonzero(d["F"])) if force else 0))
    f_fwd = 4.0 * D(s, True) + 13 * s["n"]
    if j is None:
        return f_fwd, 0.0, s["n"], 0
    return f_fwd, 4.0 * D(j, False) + 13 * j["n"] + 2 * s["n"], s["n"], j["n"]


def build_systems(w, rank=0, lo=0, hi=None):
    """systems / Jacobians [lo, hi) of the workload (rank: weak-scaling block, its own jitter stream)."""
    from pasif_b200 import workloads as wl
    n = w["n_systems"]
    hi = n if hi is None else hi
    if w.get("kind") == "duffing":
        return [wl.duffing_system()], None
    if w.get("kind") == "chain":
        s = wl.chain_system(w["ndof"], dense=w["dense"])
        if n == 1:
            return [s], [wl.chain_jacobian(s, 4)]
        # parameter replicas: the linear operator (B, K, M) is shared, every instance has its own Gamma / Lambda / F
        rng = np.random.default_rng(w.get("replica_seed", 11))
        systems = [wl.jitter_nonlinear(s, rng) for _ in range(n)][lo:hi]
        return systems, [wl.chain_jacobian(x, 4) for x in systems]
    if w["jitter_seed"] is None:
        return [wl.fit8_system() for _ in range(lo, hi)], [wl.fit8_jacobian() for _ in range(lo, hi)]
    rng = np.random.default_rng(w["jitter_seed"] + 1000 * rank)
    systems = [wl.jitter(wl.fit8_system(), rng) for _ in range(hi)][lo:]
    jacs = [wl.fit8_jacobian() for _ in range(lo, hi)]
    return systems, jacs


def build_excitations(w):
    from pasif_b200 import workloads as wl
    if w.get("kind") == "duffing":
        return np.ascontiguousarray(wl.sine_excitations(wl.DUFFING_FREQS, w["L"], 16000))
    return np.ascontiguousarray(wl.random_sine_excitations(w["E"], w["L"], 16000, seed=w["exc_seed"]))


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
        self.gpu, self.rows, self._stop, self.p = gpu, [], threading.Event(), None

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
        systems, jacs = build_systems(dict(w, n_systems=1), 0)
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
    if w.get("kind") in ("chain", "duffing") or not os.path.exists(so):
        return None
    try:
        if subprocess.run(["nvidia-smi", "-L"], capture_output=True, timeout=30).returncode != 0:
            return None
        S, E, steps = min(w["n_systems"], 2), min(w["E"], 64), 20000
        r = subprocess.run([sys.executable, os.path.join(ROOT, "oracle", "time_reference_gpu.py"), "--prec", w["dtype"],
                            "--systems", str(S), "--excitations", str(E), "--steps", str(steps), "--exc-seed", str(w["exc_seed"])],
                           capture_output=True, text=True, timeout=600)
        d = json.loads(r.stdout.strip().splitlines()[-1])
        out = {"value": d["gradient_trajectory_steps_per_s"] / w["L"], "unit": "gradient evals/s",
               "forward_system_steps_per_s": d["forward_system_steps_per_s"], "gradient_s": d["gradient_s"], "forward_s": d["forward_s"],
               "sample": f"unmodified reference (sm_100 build of src/*.cu, {w['dtype']}) through its own PASIf.py on this GPU: {S} systems x {E} "
                         f"excitations x {steps} steps (its front end cannot assemble more than 2 Jacobians; its per-step cost is "
                         f"launch-bound, i.e. independent of the batch at this size); evals/s scaled to {w['L']} steps"}
        # the headline batch itself, forward only (the one thing the reference can be timed on at this size: no Jacobians; its
        # results are invalid there -- more than 128 nonlinear non-zeros, quirk F4 -- its time is still its time)
        if w["n_systems"] > 2:
            fsteps = 4000
            r = subprocess.run([sys.executable, os.path.join(ROOT, "oracle", "time_reference_gpu.py"), "--prec", w["dtype"],
                                "--systems", str(w["n_systems"]), "--excitations", str(E), "--steps", str(fsteps), "--exc-seed", str(w["exc_seed"]),
                                "--forward-only"], capture_output=True, text=True, timeout=900)
            f = json.loads(r.stdout.strip().splitlines()[-1])
            out["forward_full_batch"] = {"system_steps_per_s": f["forward_system_steps_per_s"], "forward_s": f["forward_s"],
                                         "setSystems_s": f["setSystems_s"], "results_valid": f["results_valid"],
                                         "sample": f"{w['n_systems']} systems x {E} excitations x {fsteps} steps, getAmplitudes"}
        return out
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


# ====================================================================================================== native arm
class Ctx:
    """process-wide state of the native arm: ranks, device, peaks"""

    def __init__(self):
        import torch
        self.torch = torch
        self.rank = int(os.environ.get("RANK", "0"))
        self.local_rank = int(os.environ.get("LOCAL_RANK", "0"))
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.dist = None
        if self.world > 1:
            import torch.distributed as dist
            os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
            torch.cuda.set_device(self.local_rank)
            dist.init_process_group("nccl", device_id=torch.device("cuda", self.local_rank))
            self.dist = dist
        torch.cuda.set_device(self.local_rank)
        self.device = f"cuda:{self.local_rank}"
        self.n_sm = torch.cuda.get_device_properties(self.local_rank).multi_processor_count
        pk = os.path.join(ROOT, "MEASURED_PEAKS.json")
        self.peaks = json.load(open(pk)) if os.path.exists(pk) else {}
        self._fma = {}

    def barrier(self):
        self.torch.cuda.synchronize()
        if self.dist is not None:
            self.dist.barrier()
        self.torch.cuda.synchronize()

    def fma_peak(self, kind):
        from pasif_b200 import capi
        if kind not in self._fma:
            self._fma[kind] = capi.fma_peak(self.local_rank, kind)
        return self._fma[kind]

    def fma_roofline_peak(self, dtype, sm_mhz):
        """FMA-pipe denominator: the larger of the same-run microbenchmark (128 FMAs per loop iteration) and the
        physical peak n_sm x lanes x 2 x clock at the clock seen during the timed region."""
        lanes = 128 if dtype == "f32" else 64
        measured = max(self.fma_peak("f32"), self.fma_peak("f32x2")) if dtype == "f32" else self.fma_peak("f64")
        mhz = sm_mhz or self.peaks.get("sm_max_mhz", 1965.0)
        theo = self.n_sm * lanes * 2 * mhz * 1e6 / 1e12
        return {"peak": max(measured, theo), "peak_measured": measured, "peak_theoretical": theo,
                "peak_theoretical_formula": f"{self.n_sm} SMs x {lanes} lanes x 2 flop x {mhz:.0f} MHz",
                "peak_source": "larger of the same-run register-resident FMA microbenchmark (pasif_fma_peak, 128 FMAs per loop "
                               "iteration) and the physical pipe peak at the SM clock sampled during the timed region "
                               "(MEASURED_PEAKS.json has no FMA-pipe figure)"}

    def gather(self, vals):
        """per-rank rows of `vals` (list of floats) on every rank"""
        t = self.torch.tensor(vals, dtype=self.torch.float64, device=self.device)
        if self.dist is None:
            return [t.tolist()]
        out = [self.torch.zeros_like(t) for _ in range(self.world)]
        self.dist.all_gather(out, t)
        return [o.tolist() for o in out]


class mod:
    pass


def make_job(ctx, w, a, *, presharded, strong):
    """engine + front end for this rank's share of workload `w` through ShardedPASIf (world 1: the plain front end)."""
    from pasif_b200 import capi, frontend, workloads as wl
    from pasif_b200.sharded import ShardedPASIf, block_partition
    mod.coo_tensor = frontend.coo_tensor
    job = type("Job", (), {})()
    job.w = w
    S_total = w["n_systems"] if strong else w["n_systems"] * ctx.world
    if strong and not presharded:
        lo, hi = block_partition(w["n_systems"], ctx.world, ctx.rank) if w["n_systems"] >= ctx.world else (0, w["n_systems"])
        # every rank hands the FULL list of systems to ShardedPASIf (the public multi-GPU call); only the rank's block is
        # generated with real coefficients to keep the set-up short -- the other blocks are never read
        systems, jacs = build_systems(w, 0)
    else:
        lo, hi = 0, w["n_systems"]
        systems, jacs = build_systems(w, ctx.rank)
    job.systems, job.jacs = systems, jacs
    job.exc = build_excitations(w)
    job.forward_only = bool(w.get("forward_only")) or jacs is None
    # call_order=independent: every timed step is the same computation, the reference driver's FIRST call (the reference
    # carries forward / adjoint state from one call into the next; reproduced by the engine's default)
    opts = {"compat": a.compat, "tail_budget_gb": a.tail_budget_gb, "lanes": a.lanes, "call_order": "independent"}
    if a.lane_regs:
        opts["lane_regs"] = a.lane_regs
    if a.stage_parts:
        opts["stage_parts"] = a.stage_parts
    for kv in a.option:
        k, v = kv.split("=", 1)
        opts[k] = v
    t0 = time.time()
    sp = ShardedPASIf(job.exc, 16000, 0, driver_cls=capi.engine_class(w["dtype"]), rank=ctx.rank, world=ctx.world,
                      presharded=presharded, options=opts)
    sp.setSystems(*wl.to_frontend_args(systems, mod))
    if not job.forward_only:
        sp.setJacobian(*wl.to_frontend_args(jacs, mod, with_psi=True))
    job.setup_s = time.time() - t0
    job.sp, job.p = sp, sp.local
    lo, hi = sp._sys_range
    job.S_local = hi - lo
    job.S_total = sp.n_systems
    job.E_local = sp._exc_range[1] - sp._exc_range[0]
    job.pairs_local = job.S_local * job.E_local
    job.pairs_total = job.S_total * w["E"]
    j0 = None if job.forward_only else jacs[0]
    job.F_FWD, job.F_BWD, job.N_FWD, job.N_ADJ = system_work(systems[0], j0)
    tdt = ctx.torch.float32 if w["dtype"] == "f32" else ctx.torch.float64
    job.red = None if job.forward_only else ctx.torch.zeros(sp.n_adj_global, dtype=tdt, device=ctx.device)
    return job


def measure(ctx, w, a, steps, warmup, *, presharded=True, strong=False, headline=False, do_e2e=True):
    """Timed passes of one workload; returns the entry (rank 0) or None (other ranks)."""
    torch, dist = ctx.torch, ctx.dist
    job = make_job(ctx, w, a, presharded=presharded, strong=strong)
    p, sp = job.p, job.sp
    L, E = w["L"], w["E"]
    sz = 4 if w["dtype"] == "f32" else 8
    large = w.get("kind") == "chain"
    multi = ctx.world > 1

    def grad_step_resident():
        """(device ms of the local gradient, all-reduce ms) of one step"""
        if not multi:
            p._getGradient(0, fetch=False)
            return p._query("last_device_ms"), 0.0
        sp.getGradientReduced(device_tensor=job.red)
        return sp.last_device_ms, sp.last_allreduce_ms

    # ---------------- warm-up (>= 3), then K timed steps with resident inputs
    for _ in range(max(warmup, 3)):
        if not job.forward_only:
            grad_step_resident()
        p._getAmplitudes(fetch=False)
    sampler = ClockSampler(ctx.local_rank)
    sampler.start()
    ctx.barrier()
    wall0 = time.time()
    dev_ms = ar_ms = 0.0
    launches = 0
    per_step = []                                    # (device ms, all-reduce ms) of every timed step on this rank
    if not job.forward_only:
        for _ in range(steps):
            d_ms, r_ms = grad_step_resident()
            dev_ms += d_ms
            ar_ms += r_ms
            per_step += [d_ms, r_ms]
            launches += int(p._query("last_launches"))
    # dense-operator path: host time the last gradient spent queueing its launches (it must stay well below the device time)
    enqueue_ms = p._query("stage_enqueue_us") * 1e-3 if (not job.forward_only and _has_query(p, "stage_enqueue_us")) else 0.0
    ctx.barrier()
    wall_grad = time.time() - wall0
    # ---------------- per-kernel device times: more steps with profile=on (a CUDA-event pair around every hot-kernel
    # launch; the recompute and adjoint kernels are serialised for this pass, they overlap on two streams above)
    k_ms = {"gfwd": 0.0, "bwd": 0.0}
    k_launch = {"gfwd": 0, "bwd": 0}
    u_ms = {"upd_fwd": 0.0, "upd_adj": 0.0}          # dense-operator path only: its fused update kernels
    u_launch = {"upd_fwd": 0, "upd_adj": 0}
    serial_ms = 0.0
    prof_steps = steps if headline else min(steps, 2)
    if not job.forward_only:
        p._setOption("profile", "on")
        for _ in range(prof_steps):
            p._getGradient(0, fetch=False)
            serial_ms += p._query("last_device_ms")
            for k in k_ms:
                k_ms[k] += p._query("ms_" + k)
                k_launch[k] += int(p._query("launches_" + k))
            for k in u_ms:
                u_ms[k] += p._query("ms_" + k)
                u_launch[k] += int(p._query("launches_" + k))
        p._setOption("profile", "off")
        ctx.barrier()
    # ---------------- K forward-only passes (RK4 system-steps/s)
    fwd_ms = 0.0
    for _ in range(steps):
        p._getAmplitudes(fetch=False)
        fwd_ms += p._query("last_device_ms")
        launches += int(p._query("last_launches"))
    ctx.barrier()
    clocks = sampler.stop()
    # ---------------- e2e: host buffers in, host result out, every step
    e2e_t = 0.0
    e2e_parts = {"load_excitations_ms": 0.0, "h2d_staging_ms": 0.0, "solve_call_ms": 0.0, "device_ms": 0.0, "d2h_staging_ms": 0.0}
    e2e_steps = steps if headline else min(steps, 2)
    d2h_bytes = 0
    if do_e2e:
        exc_local = job.exc[sp._exc_range[0]:sp._exc_range[1]]
        for i in range(1 + e2e_steps):
            ctx.barrier()
            t0 = time.time()
            p._loadExcitationsSet(exc_local, 16000)
            t1 = time.time()
            if job.forward_only:
                g = p._getAmplitudes()
                d2h_bytes = g.size * sz
            elif not multi:
                g = p._getGradient(0)
                d2h_bytes = job.pairs_local * job.N_ADJ * sz
            else:
                g, red = sp.getGradientAndReduced(job.red)      # un-reduced local result to the host + all-reduce of its sum
                red_host = red.cpu()
                d2h_bytes = job.pairs_local * job.N_ADJ * sz + red_host.numel() * sz
            t2 = time.time()
            if i > 0:
                e2e_t += time.time() - t0
                e2e_parts["load_excitations_ms"] += (t1 - t0) * 1e3 / e2e_steps
                e2e_parts["solve_call_ms"] += (t2 - t1) * 1e3 / e2e_steps
                e2e_parts["device_ms"] += p._query("last_device_ms") / e2e_steps
                e2e_parts["h2d_staging_ms"] += p._query("last_h2d_ms") / e2e_steps
                e2e_parts["d2h_staging_ms"] += p._query("last_d2h_ms") / e2e_steps
    # ---------------- max over ranks; per-rank device / all-reduce times
    rows = ctx.gather([dev_ms + ar_ms, fwd_ms, e2e_t, wall_grad, dev_ms, ar_ms] + per_step)
    step_ms_tot = max(r[0] for r in rows)
    fwd_ms = max(r[1] for r in rows)
    e2e_t = max(r[2] for r in rows)
    wall_grad = max(r[3] for r in rows)
    per_rank_dev = [r[4] / max(steps, 1) for r in rows]
    per_rank_ar = [r[5] / max(steps, 1) for r in rows]
    # per step: the rank that finishes its gradient last does not wait inside the collective, so min over ranks of that
    # step's all-reduce time is the collective itself; max - min of the device times is that step's rank skew
    nccl_ms = skew_ms = 0.0
    if multi and not job.forward_only and steps:
        for k in range(steps):
            devs = [r[6 + 2 * k] for r in rows]
            ars = [r[7 + 2 * k] for r in rows]
            nccl_ms += min(ars) / steps
            skew_ms += (max(devs) - min(devs)) / steps
    # ---------------- N > 1: the all-reduced vector against a single-GPU recompute of a slice (outside the timed region)
    check = None
    if multi and not job.forward_only:
        check = allreduce_check(ctx, w, a, job, strong)
    if ctx.rank != 0:
        p.close()
        return None

    # ================================================================================================ the entry (rank 0)
    pairs = job.pairs_local
    tot_pairs = job.pairs_total
    entry = {"dtype": w["dtype"], "n_gpus": ctx.world, "steps": steps, "warmup": max(warmup, 3)}
    fwd_value = tot_pairs * L * steps / (fwd_ms * 1e-3)
    peakd = ctx.fma_roofline_peak(w["dtype"], clocks.get("sm_mhz"))
    peak = peakd["peak"]
    fwd_tf = pairs * L * job.F_FWD / (fwd_ms / steps * 1e-3) / 1e12
    entry["forward"] = {"metric": "RK4 system-steps/s", "value": fwd_value, "unit": "system-steps/s", "ms_per_step": fwd_ms / steps,
                        "roofline": {"bound": "fma", "achieved": fwd_tf, "peak": peak, "unit": "TFLOP/s", "frac": fwd_tf / peak}}
    # a batch smaller than the machine: every warp of 32 trajectories (16 packed pairs x 2 in fp32) sits alone on one SM
    # sub-partition and is bound by that partition's own FMA pipe; the chip-level fraction then mostly counts idle partitions
    lanes_per_warp = 64 if (w["dtype"] == "f32" and E % 2 == 0 and not large) else 32
    warps = -(-pairs // lanes_per_warp)
    small_batch = None
    if warps < 4 * ctx.n_sm and not large:
        small_batch = {"warps": int(warps), "sm_subpartitions": 4 * ctx.n_sm,
                       "note": "fewer warps than SM sub-partitions: fractions below are of the pipes of the sub-partitions the batch can occupy"}
    cfg = {"workload": w["name"], "systems_per_gpu": job.S_local, "systems_total": job.S_total, "excitations": E, "rk4_steps": L,
           "compat": a.compat, "sharding": ("system axis, fixed total (strong)" if strong else "system axis, per-GPU block (weak)") if multi else "single GPU",
           "l2": "working set (checkpoints + chunk trajectory) exceeds the 126 MB L2 for every throughput configuration; no flush needed",
           "timing": "CUDA events on the engine stream (+ events around the NCCL all-reduce), summed over the timed steps, max over ranks",
           "setup_s": job.setup_s, "two_trajectories_per_thread_ffma2": bool(p._query("pair_launches") > 0)}
    entry["config"] = cfg
    if small_batch:
        small_batch["forward_frac_of_occupied_subpartitions"] = entry["forward"]["roofline"]["frac"] * small_batch["sm_subpartitions"] / warps
        entry["small_batch"] = small_batch
    entry["clocks"] = clocks
    entry["gpu_launches"] = launches
    if job.forward_only:
        entry.update({"metric": "RK4 system-steps/s", "value": fwd_value, "unit": "system-steps/s", "ms_per_step": fwd_ms / steps,
                      "roofline": dict(entry["forward"]["roofline"], kernel="pasif_fwd_*", traffic=None, **peakd),
                      "roofline_hbm": {"bound": "hbm", "achieved": pairs * L * sz / (fwd_ms / steps * 1e-3) / 1e9,
                                       "peak": ctx.peaks.get("hbm_gbs", 6650.0), "unit": "GB/s",
                                       "frac": pairs * L * sz / (fwd_ms / steps * 1e-3) / 1e9 / ctx.peaks.get("hbm_gbs", 6650.0),
                                       "note": "one excitation sample per trajectory-step"}})
        if do_e2e:
            entry["e2e"] = {"value": tot_pairs * L * e2e_steps / e2e_t, "unit": "system-steps/s", "h2d_bytes_per_step": int(p._query("last_h2d_bytes")),
                            "d2h_bytes_per_step": int(d2h_bytes), "ms_per_step": e2e_t / e2e_steps * 1e3, "parts": e2e_parts}
        p.close()
        return entry

    value = tot_pairs * steps / (step_ms_tot * 1e-3)
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
    cfg["schedule"] = {"chunks": nsp, "chunk_steps": c, "chunks_kept_in_hbm": stored_chunks,
                       "forward_steps_per_eval": steps_gfwd, "adjoint_steps_per_eval": steps_bwd}
    flops_gfwd = pairs * steps_gfwd * job.F_FWD
    flops_bwd = pairs * steps_bwd * job.F_BWD
    dev_only_ms = max(per_rank_dev)                              # per step, slowest rank, without the collective
    per_kernel = {}
    for k, fl in (("gfwd", flops_gfwd), ("bwd", flops_bwd)):
        t = k_ms[k] / prof_steps * 1e-3
        per_kernel[k] = {"tflops": fl / t / 1e12 if t > 0 else None, "ms_per_step": k_ms[k] / prof_steps,
                         "launches_per_step": k_launch[k] // prof_steps,
                         "avg_launch_ms": k_ms[k] / max(k_launch[k], 1)}
    dom = max(per_kernel, key=lambda k: per_kernel[k]["ms_per_step"])
    stage = large and p._query("large_stage_classes") > 0
    lane = p._query("lane_classes") > 0
    n_needed = (job.N_FWD + 31) // 32 * 32 if large else 4
    if lane:
        # lane-block split: every lane writes / reads its share of the needed forward variables
        n_needed = int(p._query("lane_nv") * p._query("lane_G"))
    tensor = None
    if stage:
        # dense-operator path: the profiled kinds time the per-stage GEMM launches only (one GEMM per RK4 stage).
        # 3xTF32 issues three tensor-core products per useful one.
        pad = lambda n, m: (n + m - 1) // m * m  # noqa: E731
        mult = 3 if w["dtype"] == "f32" else 1
        ppad = pad(pairs, 128)
        # the GEMM covers the band of dense rows of K' (x' = v rows and probes stay row programs): rows x padded columns
        gemm_flops = {"gfwd": 2.0 * p._query("stage_band_rows_fwd") * p._query("stage_band_cols_fwd") * ppad * mult,
                      "bwd": 2.0 * p._query("stage_band_rows_adj") * p._query("stage_band_cols_adj") * ppad * mult}
        if w["dtype"] == "f32":
            tpeak = ctx.peaks.get("bf16_tflops", 1590.0) / 2.0
            tsrc = "tf32 dense = half the bf16 rate: MEASURED_PEAKS.json bf16_tflops / 2" if ctx.peaks else "fallback 1590 / 2"
            tname = ("sgemm_tn_3xtf32_tcgen05_pair_kernel (tcgen05.mma.cta_group::2 kind::tf32, TMEM accumulator, TMA operands and bulk tensor stores; "
                     "sgemm_tn_3xtf32_tcgen05_split_kernel when the shape does not pair)")
        else:
            tpeak = ctx.fma_peak("dmma")
            tsrc = "register-resident mma.sync.m8n8k4.f64 microbenchmark, same run (tcgen05 has no fp64 kind)"
            tname = "dgemm_tn_dmma_kernel<32, 2> (mma.sync.m8n8k4.f64, 32-row tiles, cp.async pipeline)"
        ach = gemm_flops[dom] / (per_kernel[dom]["avg_launch_ms"] * 1e-3) / 1e12
        whole = (flops_gfwd + flops_bwd) / (dev_only_ms * 1e-3) / 1e12
        tensor = {"bound": "tensor", "kernel": tname, "achieved": ach, "peak": tpeak, "unit": "TFLOP/s", "frac": ach / tpeak,
                  "traffic": None, "peak_source": tsrc, "tensor_products_per_useful_product": mult,
                  "useful_tflops": ach / mult, "gemm_us_per_launch": per_kernel[dom]["avg_launch_ms"] * 1e3,
                  "gemm_rows": {"fwd": p._query("stage_band_rows_fwd"), "adj": p._query("stage_band_rows_adj")},
                  "gemm_cols": {"fwd": p._query("stage_band_cols_fwd"), "adj": p._query("stage_band_cols_adj")},
                  "gemm_share_of_step": (k_ms["gfwd"] + k_ms["bwd"]) / max(serial_ms, 1e-9),
                  "whole_step": {"achieved": whole, "frac": whole / (tpeak / mult)},
                  "per_kernel": per_kernel, "serialised_ms_per_step": serial_ms / prof_steps,
                  "note": "per-GEMM time from a serialised pass (profile=on); value/whole_step from the overlapped two-part schedule; "
                          "whole_step.frac is against peak / tensor_products_per_useful_product"}
    # chunk-trajectory traffic: needed forward variables written once (recompute) and read once (adjoint)
    hbm_bytes = pairs * steps_bwd * 2 * n_needed * sz + pairs * nsp * (n_needed if large else job.N_FWD) * sz * 2
    hbm_peak = ctx.peaks.get("hbm_gbs", 6650.0)
    hbm_note = "algorithmic chunk-trajectory + checkpoint bytes of the whole gradient step over its device time"
    stage_hbm = None
    if stage:
        # dense-operator path: the state lives in HBM between the stages.  Algorithmic values moved per trajectory and
        # RK4 stage (averaged over the four stages): the GEMM reads the stage vector (band columns) and writes Y (band
        # rows); the update reads Y (band rows), Q, S (stages 2-4), the stage vector (row-program operands; just read by
        # the GEMM, so partly L2 hits) and, in the adjoint, the forward set point (twice on the half stages); it writes
        # S (stages 1-3), the next stage vector and, in stage 4, Q (+ the chunk-trajectory slot while recomputing).
        q = lambda k: p._query("stage_band_" + k)  # noqa: E731
        nrf, nra = q("nr_fwd"), q("nr_adj")
        # stage_direct (default): stage 1 evaluates at the state array (no separate stage vector to read) and stage 4 writes
        # the state array only: Q 1 + S .75 + stage vector .75 read; S .75 + stage vector .75 + Q .25 written
        per_nr = (1 + .75 + .75) + (.75 + .75 + .25) if p._query("stage_direct") else (1 + .75 + 1) + (.75 + 1 + .25)
        fused = p._query("stage_fused") > 0 if _has_query(p, "stage_fused") else False
        upd_f = q("rows_fwd") + per_nr * nrf
        upd_a = q("rows_adj") + per_nr * nra + 1.5 * nrf
        gemm_f, gemm_a = q("cols_fwd") + q("rows_fwd"), q("cols_adj") + q("rows_adj")
        if fused:                                  # Y never leaves the SM: neither written by the GEMM nor read by an update
            upd_f -= q("rows_fwd"); upd_a -= q("rows_adj"); gemm_f -= q("rows_fwd"); gemm_a -= q("rows_adj")
        stages_f, stages_a = 4 * steps_gfwd, 4 * steps_bwd
        hbm_bytes = pairs * sz * (stages_f * (upd_f + gemm_f) + (steps_gfwd - L) * nrf + stages_a * (upd_a + gemm_a))
        hbm_note = ("dense-operator path: algorithmic stage-vector bytes (GEMM operand / result + the update kernel's Y, Q, S, "
                    "stage vector, set point) of the whole gradient step over its device time")
        stage_hbm = {}
        for k, vals in (("upd_fwd", upd_f), ("upd_adj", upd_a)):
            if u_launch[k]:
                us = u_ms[k] / u_launch[k] * 1e3
                gbs = pairs * vals * sz / (us * 1e-6) / 1e9
                stage_hbm[k] = {"kernel": "stage_update_stream_kernel", "us_per_launch": us, "launches_per_step": u_launch[k] // prof_steps,
                                "ms_per_step": u_ms[k] / prof_steps, "algorithmic_bytes_per_launch": pairs * vals * sz,
                                "achieved_gbs": gbs, "frac_of_hbm_peak": gbs / hbm_peak}
    whole_tf = (flops_gfwd + flops_bwd) / (dev_only_ms * 1e-3) / 1e12
    entry.update({
        "metric": "fwd+adjoint gradient evals/s", "value": value, "unit": "gradient evals/s", "ms_per_step": step_ms_tot / steps,
        "trajectory_steps_per_s": tot_pairs * L * steps / (step_ms_tot * 1e-3),
        "roofline": dict({"bound": "fma", "kernel": ({"gfwd": "large_forward_kernel", "bwd": "large_backward_kernel"} if (large and not lane)
                                                      else {"gfwd": "pasif_gfwd_*", "bwd": "pasif_bwd_*"})[dom],
                          "achieved": per_kernel[dom]["tflops"], "unit": "TFLOP/s",
                          "frac": per_kernel[dom]["tflops"] / peak if per_kernel[dom]["tflops"] else None, "traffic": None,
                          "whole_step": {"achieved": whole_tf, "frac": whole_tf / peak},
                          "per_kernel": per_kernel,
                          "fma_peak_tflops": {k: ctx.fma_peak(k) for k in (("f32", "f32x2") if w["dtype"] == "f32" else ("f64",))},
                          "serialised_ms_per_step": serial_ms / prof_steps,
                          "note": "per-kernel figures from a serialised pass (profile=on); value/whole_step from the two-stream overlapped schedule"},
                         **peakd),
        "roofline_hbm": {"bound": "hbm", "achieved": hbm_bytes / (dev_only_ms * 1e-3) / 1e9, "peak": hbm_peak,
                         "unit": "GB/s", "frac": hbm_bytes / (dev_only_ms * 1e-3) / 1e9 / hbm_peak,
                         "note": hbm_note, "peak_source": "MEASURED_PEAKS.json" if ctx.peaks else "fallback"},
        "wall_s_grad_steps": wall_grad,
    })
    if stage and enqueue_ms > 0:
        entry["host_enqueue_ms_per_step"] = enqueue_ms
        if _has_query(p, "stage_graph_replays"):
            entry["cuda_graph_replays"] = int(p._query("stage_graph_replays"))
    if do_e2e:
        entry["e2e"] = {"value": tot_pairs * e2e_steps / e2e_t, "unit": "gradient evals/s", "h2d_bytes_per_step": int(p._query("last_h2d_bytes")),
                        "d2h_bytes_per_step": int(d2h_bytes), "steps": e2e_steps,
                        "host_buffers": "pageable float64 numpy arrays through the C ABI, staged through two pinned slots inside the engine "
                                        "(an fp32 engine narrows the samples to float while staging: half the bytes cross the link)",
                        "ms_per_step": e2e_t / e2e_steps * 1e3, "parts": e2e_parts}
    if small_batch:
        small_batch["whole_step_frac_of_occupied_subpartitions"] = entry["roofline"]["whole_step"]["frac"] * small_batch["sm_subpartitions"] / warps
    if multi:
        ar_min, ar_max = min(per_rank_ar), max(per_rank_ar)
        d_min, d_max = min(per_rank_dev), max(per_rank_dev)
        entry["allreduce_ms"] = {"collective": nccl_ms, "mean_min_over_ranks": ar_min, "mean_max_over_ranks": ar_max,
                                 "bytes": int(job.red.numel() * sz),
                                 "note": "CUDA events around dist.all_reduce on every rank and step; a rank that finishes its gradient early "
                                         "waits inside the collective for the slowest rank, so `collective` = mean over steps of the minimum "
                                         "over ranks is the all-reduce itself; the per-rank means include that waiting"}
        entry["device_ms_per_rank"] = {"min": d_min, "max": d_max, "all": per_rank_dev}
        entry["scaling_loss"] = {"rank_skew_ms": skew_ms, "nccl_ms": nccl_ms,
                                 "note": "per step: max - min over ranks of the local gradient's device time (skew) against the "
                                         "collective's own time; means over the timed steps",
                                 "limiter": "rank skew (kernel time differs between GPUs and steps)" if skew_ms > nccl_ms else "NCCL all-reduce"}
        entry["allreduce_check"] = check
    if lane:
        cfg["lane_split"] = {k: p._query("lane_" + k) for k in ("G", "R", "RA", "nv", "nc_fwd", "nc_adj", "halo_fwd", "halo_adj", "halo_p", "efficiency")}
        cfg["path"] = "lane-block register kernels (lane_split.hpp): G threads per trajectory, state in registers, halo exchange per stage"
    elif large:
        cfg["path"] = "dense-operator stage path (tensor-core GEMM per stage)" if stage else "shared-memory tile kernels"
    else:
        cfg["path"] = "register kernels, one thread per " + ("two trajectories (packed FFMA2)" if cfg["two_trajectories_per_thread_ffma2"] else "trajectory")
    if stage_hbm:
        entry["roofline_hbm"]["update_kernels"] = stage_hbm
    if tensor is not None:
        entry["roofline_fma"] = entry["roofline"]
        entry["roofline"] = tensor
        entry["forward"]["roofline"] = {"bound": "tensor", "achieved": fwd_tf, "unit": "TFLOP/s",
                                        "peak": tensor["peak"] / tensor["tensor_products_per_useful_product"],
                                        "frac": fwd_tf / (tensor["peak"] / tensor["tensor_products_per_useful_product"])}
    p.close()
    return entry


def _has_query(p, key):
    try:
        p._query(key)
        return True
    except KeyError:
        return False


def allreduce_check(ctx, w, a, job, strong, n_check=32):
    """N > 1, outside the timed region: rank 0 recomputes the excitation-summed gradient of the FIRST `n_check` systems of
    the LAST rank's block on its own GPU with a fresh single-GPU engine and compares it with that slice of the all-reduced
    vector (fp64: <= 1e-12; the kernels are deterministic and the all-reduce only adds zeros, so the expected error is 0)."""
    from pasif_b200 import capi, frontend, workloads as wl
    from pasif_b200.sharded import block_partition
    sp = job.sp
    sp.getGradientReduced(device_tensor=job.red)                 # a fresh all-reduced vector on every rank
    out = None
    if ctx.rank == 0:
        last = ctx.world - 1
        if strong:
            lo, hi = block_partition(w["n_systems"], ctx.world, last)
            systems, jacs = build_systems(w, 0, lo, min(hi, lo + n_check))
            first_sys = lo
        else:
            systems, jacs = build_systems(w, last, 0, min(w["n_systems"], n_check))
            first_sys = last * w["n_systems"]
        n = len(systems)
        q = frontend.make_frontend(capi.engine_class(w["dtype"]))(job.exc, 16000, 0)
        q._setOption("compat", a.compat)
        q._setOption("call_order", "independent")
        q.setSystems(*wl.to_frontend_args(systems, mod))
        q.setJacobian(*wl.to_frontend_args(jacs, mod, with_psi=True))
        ref = np.asarray(q._getGradientReduced(0), dtype=np.float64)
        q.close()
        off = first_sys * job.N_ADJ
        got = job.red[off:off + n * job.N_ADJ].double().cpu().numpy()
        err = float(np.max(np.abs(got - ref)) / max(np.max(np.abs(ref)), 1e-300))
        tol = 1e-12 if w["dtype"] == "f64" else 1e-5
        out = {"rel_err": err, "tolerance": tol, "ok": bool(err <= tol), "systems_checked": n, "first_system": first_sys,
               "what": f"all-reduced [S x N_adj] vector, slice of rank {last}'s block, against a single-GPU recompute on rank 0"}
        assert err <= tol, f"all-reduced gradient differs from the single-GPU recompute: rel err {err:.3e} > {tol}"
    ctx.barrier()
    return out


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
    ap.add_argument("--configs", default="auto", choices=["auto", "all", "none"],
                    help="the per-config entries of the other BASELINE configs (auto: default workload on one GPU, no overrides)")
    ap.add_argument("--c5-strong", default="auto", choices=["auto", "on", "off"],
                    help="BASELINE config 5 at its full size, fixed total sharded over the ranks (auto: default workload, no overrides)")
    ap.add_argument("--aux-steps", type=int, default=3, help="timed steps of the per-config entries and of c5_strong")
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
    overridden = bool(a.systems or a.length or a.excitations or a.dtype)
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

    ctx = Ctx()
    t_run = time.time()
    default_run = a.workload == "c3" and not overridden
    strong = bool(w.get("strong"))
    line = measure(ctx, w, a, a.steps, a.warmup, presharded=not strong, strong=strong, headline=True)
    if ctx.rank == 0:
        line["scaling"] = "strong" if strong else "weak"
        line["higher_is_better"] = True
        line["vs_baseline"] = None
        line["data"] = "synthetic"
        tr = NCU_TRAFFIC.get((a.workload, w["dtype"], a.compat))
        if tr and not overridden and "per_kernel" in line.get("roofline", {}):
            dom = "gfwd" if line["roofline"]["kernel"].startswith("pasif_gfwd") else "bwd"
            nsp, c = schedule(w["L"], w["dtype"])
            line["roofline"]["traffic"] = tr[dom]
            line["roofline"]["traffic_source"] = tr["source"] + " (ncu --set full, same command; bytes per launch of the dominant kernel)"
            line["roofline"]["algorithmic_bytes_per_launch"] = w["n_systems"] * w["E"] * c * 4 * (4 if w["dtype"] == "f32" else 8)

    # ---------------- BASELINE config 5 at full size, fixed total over the ranks
    if (a.c5_strong == "on") or (a.c5_strong == "auto" and default_run):
        e = measure(ctx, dict(WORKLOADS["c5s"]), a, a.aux_steps, 3, presharded=False, strong=True)
        if ctx.rank == 0:
            e["scaling"] = "strong"
            line["c5_strong"] = e

    # ---------------- the other BASELINE configs (one GPU)
    if ctx.world == 1 and ((a.configs == "all") or (a.configs == "auto" and default_run)):
        configs = {}
        plan = [("c1", "c1", {}), ("c2", "c2", {}),
                ("c4_sparse_f32", "c4s", {"dtype": "f32"}), ("c4_sparse_f64", "c4s", {"dtype": "f64"}),
                ("c4_dense_f32", "c4d", {"dtype": "f32"}), ("c4_dense_f32_512steps", "c4d", {"dtype": "f32", "L": 512}),
                ("c4_dense_f64", "c4d", {"dtype": "f64"}), ("c4_dense_f64_512steps", "c4d", {"dtype": "f64", "L": 512}),
                ("c4_replicas_f32", "c4r", {}),
                # exact schedule (compat=clean): the forward states of all 128 steps stay in HBM (4.6 GB), one sweep, no recompute
                ("c4_dense_f32_exact_schedule", "c4d", {"dtype": "f32", "compat": "clean"}),
                # the headline workload and C4 sparse in the exact schedule: chunk trajectories of the tail chunks stay in free HBM
                ("c3_exact_schedule", "c3", {"compat": "clean"}),
                ("c4_sparse_f32_exact_schedule", "c4s", {"dtype": "f32", "compat": "clean"})]
        for key, wk, over in plan:
            ww = dict(WORKLOADS[wk], **{k: v for k, v in over.items() if k != "compat"})
            aa = argparse.Namespace(**dict(vars(a), compat=over["compat"])) if "compat" in over else a
            if "L" in over:
                ww["name"] = ww["name"].replace("x 128 RK4 steps", f"x {over['L']} RK4 steps")
            if "dtype" in over:
                ww["name"] += f" [{over['dtype']}]"
            try:
                configs[key] = measure(ctx, ww, aa, a.aux_steps, 3)
            except Exception as ex:  # a failing side entry must not take the headline down with it
                configs[key] = {"error": repr(ex)[:300]}
        # the dense-operator rate per trajectory-step at the two lengths (step-count independence)
        for dt in ("f32", "f64"):
            x, y = configs.get(f"c4_dense_{dt}"), configs.get(f"c4_dense_{dt}_512steps")
            if x and y and "trajectory_steps_per_s" in x and "trajectory_steps_per_s" in y:
                y["rate_vs_128_steps"] = y["trajectory_steps_per_s"] / x["trajectory_steps_per_s"]
        line["configs"] = configs

    if ctx.rank != 0:
        if ctx.dist is not None:
            ctx.dist.destroy_process_group()
        return
    if not a.no_cpu_baseline and ctx.world == 1 and not w.get("forward_only"):
        threads = os.cpu_count() or 1
        v, fv, sample, _ = cpu_baseline_sample(w, threads)
        line["cpu_baseline"] = {"value": v, "unit": "gradient evals/s", "cores": threads, "kind": "port", "sample": sample,
                                "forward_system_steps_per_s": fv}
    line["bench_wall_s"] = time.time() - t_run
    # key order of the contract first
    head = ["metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline", "dtype",
            "data", "config"]
    out = {k: line[k] for k in head if k in line}
    out.update({k: v for k, v in line.items() if k not in out})
    sys.stdout.flush()
    os.write(real_stdout, (json.dumps(out) + "\n").encode())
    if ctx.dist is not None:
        ctx.dist.destroy_process_group()


if __name__ == "__main__":
    main()
