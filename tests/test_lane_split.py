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
