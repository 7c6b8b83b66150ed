"""GPU suite: the N > 1 path (pasif_b200/sharded.py) on real devices -- the DEFAULT driver (pybind module PASIfgpu),
the on-device reduced gradient written into a torch tensor, and one NCCL all-reduce -- against the single-GPU answer
and the C oracle.  The two-rank test needs >= 2 GPUs (skipped on a one-GPU box); the world-size-1 test runs anywhere."""
import os
import socket
import subprocess
import sys

import numpy as np
import pytest

from conftest import ROOT, frontend_mod, rel_err

pytestmark = pytest.mark.gpu

WORKER = r"""
import os, sys
import numpy as np
sys.path[:0] = [%(root)r, %(root)r + "/oracle"]
import torch
import torch.distributed as dist
from pasif_b200 import frontend, workloads as wl
from pasif_b200.sharded import ShardedPASIf
class mod: coo_tensor = frontend.coo_tensor
rank = int(os.environ["RANK"])
torch.cuda.set_device(rank)
dist.init_process_group("nccl", init_method="tcp://127.0.0.1:%(port)d", rank=rank, world_size=2, device_id=torch.device("cuda", rank))
rng = np.random.default_rng(3)
out = {}
for axis, n_sys, E in (("excitations", 1, 6), ("systems", 5, 6), ("systems", 64, 32)):
    systems = [wl.jitter(wl.fit8_system(), rng) for _ in range(n_sys)]
    jacs = [wl.jitter(wl.fit8_jacobian(), rng) for _ in range(n_sys)]
    exc = wl.random_sine_excitations(E, 700, 16000, seed=9)
    sp = ShardedPASIf([np.array(r) for r in exc], 16000, 0)              # default driver: the pybind module, fp64
    sp.setSystems(*wl.to_frontend_args(systems, mod))
    sp.setJacobian(*wl.to_frontend_args(jacs, mod, with_psi=True))
    assert sp._axis == axis, (sp._axis, axis)
    red = torch.full((sp.n_adj_global,), 7.0, dtype=torch.float64, device=f"cuda:{rank}")   # stale contents must not leak
    sp.getGradientReduced(device_tensor=red)
    host = sp.getGradientReduced()                                       # host path, same collective on a CPU tensor ...
    loc, red2 = sp.getGradientAndReduced(torch.empty_like(red))          # ... and both results from one evaluation
    assert np.array_equal(red.cpu().numpy(), red2.cpu().numpy())
    # single-GPU answer on this rank's device
    one = frontend.make_frontend(frontend.load_native_driver())([np.array(r) for r in exc], 16000, 0)
    one.setSystems(*wl.to_frontend_args(systems, mod))
    one.setJacobian(*wl.to_frontend_args(jacs, mod, with_psi=True))
    full = np.asarray(one.getGradient()).reshape(E, -1)
    ref = full.sum(axis=0)
    err = float(np.max(np.abs(red.cpu().numpy() - ref)) / np.max(np.abs(ref)))
    assert err < 1e-12, (axis, err)
    if axis == "systems":                                                # disjoint blocks: the all-reduce only adds zeros
        lo, hi = sp.adj_slice
        assert np.array_equal(np.asarray(loc).reshape(E, -1), full[:, lo:hi])
    out[(axis, n_sys)] = err
if rank == 0:
    print("SHARDED_GPU_OK", out)
dist.destroy_process_group()
"""


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _n_gpus():
    import torch
    return torch.cuda.device_count()


@pytest.mark.skipif("_n_gpus() < 2", reason="needs two GPUs")
def test_sharded_gradient_allreduce_nccl_world2(tmp_path):
    script = tmp_path / "worker.py"
    script.write_text(WORKER % dict(root=ROOT, port=_free_port()))
    procs = []
    for rank in range(2):
        env = dict(os.environ, RANK=str(rank), LOCAL_RANK=str(rank), WORLD_SIZE="2", MASTER_ADDR="127.0.0.1")
        env.pop("PASIF_DTYPE", None)
        procs.append(subprocess.Popen([sys.executable, str(script)], env=env, stdout=subprocess.PIPE,
                                      stderr=subprocess.STDOUT, text=True))
    outs = [p.communicate(timeout=600)[0] for p in procs]
    assert all(p.returncode == 0 for p in procs), "\n".join(outs)
    assert "SHARDED_GPU_OK" in outs[0]


def test_device_tensor_path_with_default_driver_world1():
    """ShardedPASIf on one rank, default (pybind) driver: the reduced gradient lands in a torch tensor on the device and
    equals the host path and the oracle's sum over excitations."""
    import torch
    from pasif_oracle_c import CRecordingDriver
    from pasif_b200 import frontend, workloads as wl
    from pasif_b200.sharded import ShardedPASIf
    rng = np.random.default_rng(5)
    systems = [wl.jitter(wl.fit8_system(), rng) for _ in range(3)]
    jacs = [wl.jitter(wl.fit8_jacobian(), rng) for _ in range(3)]
    exc = [np.array(r) for r in wl.random_sine_excitations(10, 500, 16000, seed=4)]
    sp = ShardedPASIf(exc, 16000, 0, rank=0, world=1)
    sp.setSystems(*wl.to_frontend_args(systems, frontend_mod))
    sp.setJacobian(*wl.to_frontend_args(jacs, frontend_mod, with_psi=True))
    red = torch.full((sp.n_adj_global,), -3.0, dtype=torch.float64, device="cuda:0")
    sp.getGradientReduced(device_tensor=red)
    assert sp.last_device_ms > 0 and sp.last_allreduce_ms == 0.0
    host = sp.getGradientReduced()
    assert np.array_equal(red.cpu().numpy(), host)
    loc, red2 = sp.getGradientAndReduced(torch.empty_like(red))
    assert np.array_equal(red2.cpu().numpy(), host)
    orc = frontend.make_frontend(CRecordingDriver)(exc, 16000, 0)
    orc.setSystems(*wl.to_frontend_args(systems, frontend_mod))
    orc.setJacobian(*wl.to_frontend_args(jacs, frontend_mod, with_psi=True))
    g = np.asarray(orc.getGradient()).reshape(10, -1)
    assert rel_err(np.asarray(loc).reshape(10, -1), g) < 1e-12
    assert rel_err(host, g.sum(axis=0)) < 1e-12
    with pytest.raises(ValueError):
        sp.getGradientReduced(device_tensor=torch.zeros(3, dtype=torch.float64, device="cuda:0"))
