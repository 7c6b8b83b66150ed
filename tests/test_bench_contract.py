"""CPU suite: the parts of bench.py's contract that need no GPU -- the reference arm's JSON line (a bounded sample of the
workload through the C oracle on the host cores) and the native arm failing loudly, not falling back, without a GPU."""
import json
import os
import subprocess
import sys

import pytest

from conftest import ROOT


def _has_gpu():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


def _run(args, env=None):
    return subprocess.run([sys.executable, os.path.join(ROOT, "bench.py")] + args, stdout=subprocess.PIPE, stderr=subprocess.PIPE,
                          text=True, env=env, timeout=600)


def test_reference_arm_line():
    r = _run(["--impl", "reference", "--steps", "1", "--warmup", "0", "--length", "2000"])
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.strip()]
    assert len(lines) == 1                                   # stdout carries the JSON line and nothing else
    d = json.loads(lines[0])
    assert d["impl"] == "reference"
    assert d["metric"] == "fwd+adjoint gradient evals/s" and d["unit"] == "gradient evals/s"
    assert d["higher_is_better"] is True and d["n_gpus"] == 1 and d["steps"] == 1
    assert d["value"] > 0 and d["ms_per_step"] > 0
    assert d["config"]["workload"].startswith("C3")
    cb = d["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] >= 1 and cb["value"] == d["value"] and "steps" in cb["sample"]
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert d["gpu_launches"] == 0


def test_reference_arm_other_ranks_exit_quietly():
    env = dict(os.environ, RANK="1", LOCAL_RANK="1", WORLD_SIZE="2")
    r = _run(["--impl", "reference", "--gpus", "2", "--steps", "1", "--warmup", "0", "--length", "2000"], env=env)
    assert r.returncode == 0 and r.stdout.strip() == ""


@pytest.mark.skipif(_has_gpu(), reason="only meaningful on a box without a GPU")
def test_native_arm_has_no_cpu_fallback():
    r = _run(["--steps", "1", "--warmup", "1", "--systems", "4", "--length", "100"])
    assert r.returncode != 0
    assert r.stdout.strip() == ""                            # no bench line is printed for work that did not run on a GPU


def test_bench_host_logic():
    """The pieces of bench.py that decide WHAT is measured: workload slices are reproducible (the all-reduce check on rank 0
    rebuilds systems of another rank's block), the literal schedule matches the engine's, the flop model matches DESIGN.md."""
    import numpy as np
    sys.path.insert(0, ROOT)
    import bench
    w = dict(bench.WORKLOADS["c5s"], n_systems=40)
    full, jf = bench.build_systems(w, 0)
    part, jp = bench.build_systems(w, 0, 25, 33)
    assert len(part) == 8 and len(jp) == 8
    for a, b in zip(full[25:33], part):
        assert np.array_equal(a["K"][0], b["K"][0]) and np.array_equal(a["F"], b["F"]) and a["Gamma"][1] == b["Gamma"][1]
    other, _ = bench.build_systems(w, 3, 0, 4)                      # another rank's jitter stream differs
    assert not np.array_equal(other[0]["K"][0], full[0]["K"][0])
    assert bench.system_work(full[0], jf[0]) == (240.0, 380.0, 8, 12)
    assert bench.schedule(100000, "f32") == (316, 317) and bench.schedule(78001, "f64") == (279, 280)
    c1, j1 = bench.build_systems(bench.WORKLOADS["c1"], 0)
    assert j1 is None and bench.system_work(c1[0], None)[0] == 120.0
    r, jr = bench.build_systems(dict(bench.WORKLOADS["c4r"], n_systems=3, ndof=6), 0)
    assert len(r) == 3 and np.array_equal(r[0]["K"][0], r[1]["K"][0]) and r[0]["Lambda"][1] != r[1]["Lambda"][1]
    assert set(bench.WORKLOADS) >= {"c1", "c2", "c3", "c4s", "c4d", "c4r", "c5", "c5s"}
