"""CPU suite: the N > 1 path (pasif_b200/sharded.py) with world_size 2 over gloo, the oracle standing in for the device."""
import os
import socket
import subprocess
import sys

import numpy as np

from conftest import ROOT
from pasif_b200.sharded import block_partition

WORKER = r"""
import os, sys
import numpy as np
sys.path[:0] = [%(root)r, %(root)r + "/oracle"]
import torch.distributed as dist
from pasif_oracle_c import CRecordingDriver
from pasif_b200 import frontend, workloads as wl
from pasif_b200.sharded import ShardedPASIf
class mod: coo_tensor = frontend.coo_tensor
dist.init_process_group("gloo", init_method="tcp://127.0.0.1:%(port)d", rank=int(os.environ["RANK"]), world_size=2)
rng = np.random.default_rng(3)
out = {}
for axis, n_sys in (("excitations", 1), ("systems", 4)):
    systems = [wl.jitter(wl.fit8_system(), rng) for _ in range(n_sys)]
    jacs = [wl.jitter(wl.fit8_jacobian(), rng) for _ in range(n_sys)]
    exc = wl.random_sine_excitations(6, 300, 16000, seed=9)
    sp = ShardedPASIf([np.array(r) for r in exc], 16000, 0, driver_cls=CRecordingDriver)
    sp.setSystems(*wl.to_frontend_args(systems, mod))
    sp.setJacobian(*wl.to_frontend_args(jacs, mod, with_psi=True))
    assert sp._axis == axis, (sp._axis, axis)
    red = sp.getGradientReduced()
    # single-process answer
    one = frontend.make_frontend(CRecordingDriver)([np.array(r) for r in exc], 16000, 0)
    one.setSystems(*wl.to_frontend_args(systems, mod))
    one.setJacobian(*wl.to_frontend_args(jacs, mod, with_psi=True))
    full = np.asarray(one.getGradient()).reshape(6, -1).sum(axis=0)
    err = float(np.max(np.abs(red - full)) / np.max(np.abs(full)))
    assert err < 1e-13, (axis, err)
    out[axis] = err
# presharded: every rank hands over only its own block of systems (weak scaling); offsets come from an all-gather
rank = dist.get_rank()
blocks = [[wl.jitter(wl.fit8_system(), np.random.default_rng(50 + r)) for _ in range(2 + r)] for r in range(2)]
jblocks = [[wl.jitter(wl.fit8_jacobian(), np.random.default_rng(60 + r)) for _ in range(2 + r)] for r in range(2)]
exc = wl.random_sine_excitations(4, 250, 16000, seed=3)
sp = ShardedPASIf([np.array(r) for r in exc], 16000, 0, driver_cls=CRecordingDriver, presharded=True)
sp.setSystems(*wl.to_frontend_args(blocks[rank], mod))
sp.setJacobian(*wl.to_frontend_args(jblocks[rank], mod, with_psi=True))
assert sp.n_systems == 5 and sp._sys_range == ((0, 2) if rank == 0 else (2, 5)) and sp.n_adj_global == 60
assert sp.adj_slice == ((0, 24) if rank == 0 else (24, 60))
red = sp.getGradientReduced()
one = frontend.make_frontend(CRecordingDriver)([np.array(r) for r in exc], 16000, 0)
one.setSystems(*wl.to_frontend_args(blocks[0] + blocks[1], mod))
one.setJacobian(*wl.to_frontend_args(jblocks[0] + jblocks[1], mod, with_psi=True))
full = np.asarray(one.getGradient()).reshape(4, -1).sum(axis=0)
err = float(np.max(np.abs(red - full)) / np.max(np.abs(full)))
assert err < 1e-13, ("presharded", err)
out["presharded"] = err
if dist.get_rank() == 0:
    print("SHARDED_OK", out)
dist.destroy_process_group()
"""


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def test_block_partition_covers_everything():
    for n in (1, 7, 64, 1000):
        for w in (1, 2, 3, 8):
            parts = [block_partition(n, w, r) for r in range(w)]
            assert parts[0][0] == 0 and parts[-1][1] == n
            assert all(parts[i][1] == parts[i + 1][0] for i in range(w - 1))
            sizes = [b - a for a, b in parts]
            assert max(sizes) - min(sizes) <= 1


def test_choose_axis():
    from pasif_b200.sharded import choose_axis
    import pytest
    assert choose_axis(8, 2, 8) == "systems" and choose_axis(1, 64, 8) == "excitations" and choose_axis(3, 64, 2) == "systems"
    with pytest.raises(ValueError):
        choose_axis(3, 5, 8)


def test_sharded_gradient_allreduce_gloo_world2(tmp_path):
    script = tmp_path / "worker.py"
    script.write_text(WORKER % dict(root=ROOT, port=_free_port()))
    procs = []
    for rank in range(2):
        env = dict(os.environ, RANK=str(rank), WORLD_SIZE="2", MASTER_ADDR="127.0.0.1", OMP_NUM_THREADS="2")
        procs.append(subprocess.Popen([sys.executable, str(script)], env=env, stdout=subprocess.PIPE,
                                      stderr=subprocess.STDOUT, text=True))
    outs = [p.communicate(timeout=300)[0] for p in procs]
    assert all(p.returncode == 0 for p in procs), "\n".join(outs)
    assert "SHARDED_OK" in outs[0]
