"""CPU suite for the lane-block partitioner (pasif_b200/csrc/lane_split.hpp) through the C ABI on a host-only engine:
the split runs its own self-check (one derivative evaluation of the source pattern against the lane pattern, per side)
and raises if it fails, so a successful `setSystems` / `setJacobian` already pins the tables; here the decisions of the
cost model are checked, and that the specialised kernels compile for sm_100a without spills (NVRTC needs no GPU)."""
import re

import numpy as np
import pytest

from pasif_b200 import capi, frontend, workloads as wl


class _mod:
    coo_tensor = frontend.coo_tensor


def _host(prec, systems, jacobians, options=None):
    P = frontend.make_frontend(capi.host_engine_class(prec))
    p = P([np.zeros(8)] * 3, 16000, 0)
    for k, v in (options or {}).items():
        p._setOption(k, v)
    p.setSystems(*wl.to_frontend_args(systems, _mod))
    if jacobians is not None:
        p.setJacobian(*wl.to_frontend_args(jacobians, _mod, with_psi=True))
    return p


@pytest.mark.parametrize("prec,ndof,G,eff", [("f32", 256, 32, 0.9), ("f64", 256, 64, 0.85), ("f32", 40, 8, 0.85),
                                              ("f64", 100, 32, 0.65)])        # 100 units on 32 x 4 positions: 22 % idle
def test_chain_is_split_over_lane_groups(prec, ndof, G, eff):
    s = wl.chain_system(ndof)
    p = _host(prec, [s], [wl.chain_jacobian(s, 4)])
    assert p._query("lane_classes") == 2 and p._query("large_classes") == 0
    assert p._query("lane_selfcheck_failures") == 0
    assert p._query("lane_G") == G
    assert p._query("lane_halo_fwd") == 2 and p._query("lane_halo_adj") == 2 and p._query("lane_halo_p") == 0
    assert p._query("lane_efficiency") > eff
    assert p._query("joint_instances") == G                       # virtual instances of the one real instance


def test_membrane_is_split_with_row_halo():
    s = wl.lattice_system(16, 16)
    p = _host("f32", [s], [wl.chain_jacobian(s, 3)])
    assert p._query("lane_classes") == 2 and p._query("lane_G") == 32
    assert p._query("lane_halo_fwd") == 18 and p._query("lane_efficiency") > 0.9


def test_dense_operator_and_small_systems_are_not_split():
    s = wl.chain_system(64, dense=True)
    p = _host("f32", [s], None)
    assert p._query("lane_classes") == 0                           # left to the dense-operator (tensor-core) path
    p = _host("f64", [wl.fit8_system()], [wl.fit8_jacobian()])
    assert p._query("lane_classes") == 0                           # register-resident: one thread per trajectory
    s = wl.chain_system(16)                                        # 34 states: fits one thread in fp32, would spill in fp64
    assert _host("f32", [s], [wl.chain_jacobian(s, 3)])._query("lane_classes") == 0
    assert _host("f64", [s], [wl.chain_jacobian(s, 3)])._query("lane_classes") == 2
    s = wl.chain_system(40)
    p = _host("f32", [s], None, {"lanes": "off"})
    assert p._query("lane_classes") == 0


@pytest.mark.parametrize("prec", ["f32", "f64"])
def test_lane_kernels_compile_for_sm100a_without_spills(prec):
    s = wl.chain_system(40)
    p = _host(prec, [s], [wl.chain_jacobian(s, 3)])
    log = p.jit_log("sm_100a")
    assert log.count("rc=0") == 2 and "rc=1" not in log
    assert len(re.findall(r"Compiling entry function 'pasif_k_(?:fwd|gfwd|bwd)'", log)) == 3
    assert all(int(x) == 0 for x in re.findall(r"(\d+) bytes spill stores", log))
    assert all(int(x) <= 255 for x in re.findall(r"Used (\d+) registers", log))


def _random_system(rng, n, band, extra):
    """first-order system with a banded linear part plus `extra` random long-range couplings and random quadratic / cubic
    terms: not a mechanical chain -- the partitioner must either split it correctly or refuse it"""
    from pasif_b200.workloads import _sys
    kv, kr, kc = [], [], []
    for i in range(n):
        for d in range(-band, band + 1):
            j = i + d
            if 0 <= j < n and (d == 0 or rng.random() < 0.7):
                kv.append(rng.uniform(-2, 2)); kr.append(i); kc.append(j)
    for _ in range(extra):
        i, j = rng.integers(0, n, size=2)
        kv.append(rng.uniform(-1, 1)); kr.append(int(i)); kc.append(int(j))
    # scipy sums duplicates on conversion; keep (row, col) unique
    seen, kv2, kr2, kc2 = set(), [], [], []
    for v, r, c in zip(kv, kr, kc):
        if (r, c) not in seen:
            seen.add((r, c)); kv2.append(v); kr2.append(r); kc2.append(c)
    gval, gidx, lval, lidx = [], [], [], []
    for _ in range(n // 2):
        o, a = rng.integers(0, n, size=2)
        b = min(n - 1, max(0, int(a) + int(rng.integers(-2, 3))))
        gval.append(rng.uniform(-1, 1)); gidx += [int(o), int(a), b]
    for _ in range(n // 3):
        o = int(rng.integers(0, n))
        lval.append(rng.uniform(-1, 1)); lidx += [o, o, max(0, o - 1), min(n - 1, o + 1)]
    F = np.where(rng.random(n) < 0.3, rng.uniform(0.5, 1.5, size=n), 0.0)
    empty = (np.zeros(0), np.zeros(0, np.int64), np.zeros(0, np.int64))
    return _sys(n, np.ones(n), empty, (kv2, kr2, kc2), ([n] * 3, gval, gidx), ([n] * 4, lval, lidx), F, rng.uniform(-1, 1, size=n))


@pytest.mark.parametrize("seed", range(8))
def test_random_sparse_systems_never_fail_the_self_check(seed):
    """whatever the cost model decides, a split that is accepted has passed the derivative self-check on the host, and a
    system that is refused stays with the other paths -- no exception, no failed check"""
    from pasif_b200 import adjoint
    rng = np.random.default_rng(100 + seed)
    n = int(rng.integers(50, 160))
    s = _random_system(rng, n, band=int(rng.integers(1, 4)), extra=int(rng.integers(0, 6)))
    loss = np.zeros(n); loss[int(rng.integers(0, n))] = 1.0
    nl = len(s["Lambda"][1])
    j = adjoint.build_adjoint(s, loss, [("Lambda", 0), ("Gamma", 1), ("K", 0, 0)] if nl else [("Gamma", 0)])
    for prec in ("f32", "f64"):
        p = _host(prec, [s], [j])
        assert p._query("lane_selfcheck_failures") == 0
        assert p._query("joint_instances") >= 1
    # with the cost model switched off the irregular system IS split (a poor split: mostly padding terms) and the tables
    # -- halo entries at arbitrary lane offsets, adjoint rows attached by affinity -- still reproduce the derivative
    p = _host("f64", [s], [j], {"lane_regs": "1000000", "lane_min_eff": "0", "lanes": "16"})
    assert p._query("lane_selfcheck_failures") == 0 and p._query("lane_classes") >= 1
