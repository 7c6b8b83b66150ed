"""GPU suite for the LANE-BLOCK path (pasif_b200/csrc/lane_split.hpp + the lane-aware register kernels in
rk4_kernels.cuh): a component too large for one thread is integrated by a group of G threads, every lane keeping its
share of the state in registers and exchanging a few halo operands per RK4 stage.  Called through the C ABI and compared
with the C oracle (the restated reference algorithm) and with the tile kernels on the same inputs:
  * sparse oscillator chains (BASELINE config 4 shape, scaled down), forward + trajectory + adjoint, both schedules, both
    precisions, several group widths (lanes = 8 .. 128, one warp and several warps per trajectory);
  * several instances with their own coefficients; an excitation count that leaves the last CTA partly filled;
  * the reference's golden vectors with every component forced through the split (whatever the cost model then keeps).
Tolerances as in test_parity_gpu.py (fp64 <= 1e-12 per variable; fp32 <= 1e-5 forward / 5e-5 adjoint on the vector)."""
import numpy as np
import pytest

from conftest import golden, rel_err, col_err, run_case
from pasif_b200 import capi, workloads as wl
from test_large_gpu import TOL_ADJ, TOL_FWD, _oracle, chain_case, close

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("prec", ["f64", "f32"])
@pytest.mark.parametrize("compat", ["reference", "clean"])
@pytest.mark.parametrize("ndof,lanes", [(40, "on"), (40, "16"), (100, "on"), (150, "128")])
def test_chain_through_lane_groups_matches_oracle(ndof, lanes, compat, prec):
    case = chain_case(ndof, False, E=11, L=700)
    a = run_case(capi.engine_class(prec, compat), case, options={"lanes": lanes})
    b = run_case(_oracle(prec, compat), case)
    p = a["_p"]
    assert p._query("lane_classes") >= 2 and p._query("large_classes") == 0      # forward class + joint class
    assert p._query("lane_G") <= (128 if lanes == "on" else int(lanes))
    n, na = case["systems"][0]["n"], case["jacobians"][0]["n"]
    assert close(a["amplitudes"], b["amplitudes"], n, prec, TOL_FWD)
    assert np.array_equal(a["amplitudes"], a["amplitudes_2nd"])
    assert rel_err(a["trajectory"], b["trajectory"]) < TOL_FWD[prec]
    assert close(a["gradient"], b["gradient"], na, prec, TOL_ADJ)


@pytest.mark.parametrize("prec", ["f64", "f32"])
def test_lane_groups_agree_with_tile_kernels(prec):
    """same inputs through the lane-block split and through the shared-memory tile kernels"""
    case = chain_case(60, False, E=37, L=400)
    a = run_case(capi.engine_class(prec), case)
    t = run_case(capi.engine_class(prec), case, options={"lanes": "off"})
    assert a["_p"]._query("lane_classes") >= 2 and t["_p"]._query("lane_classes") == 0
    n, na = case["systems"][0]["n"], case["jacobians"][0]["n"]
    assert close(a["amplitudes"], t["amplitudes"], n, prec, TOL_FWD)
    assert close(a["gradient"], t["gradient"], na, prec, TOL_ADJ)


@pytest.mark.parametrize("prec", ["f64", "f32"])
def test_several_chain_instances_through_lane_groups(prec):
    case = chain_case(30, False, E=19, L=500, n_sys=3)
    a = run_case(capi.engine_class(prec), case, options={"call_order": "independent"})   # two gradient calls below
    b = run_case(_oracle(prec, "reference"), case)
    assert a["_p"]._query("lane_classes") >= 2
    n = sum(s["n"] for s in case["systems"])
    na = sum(s["n"] for s in case["jacobians"])
    assert close(a["amplitudes"], b["amplitudes"], n, prec, TOL_FWD)
    assert close(a["gradient"], b["gradient"], na, prec, TOL_ADJ)
    red = np.asarray(a["_p"]._getGradientReduced(0))
    assert rel_err(red, a["gradient"].reshape(19, na).sum(axis=0)) < (1e-12 if prec == "f64" else 1e-5)


@pytest.mark.parametrize("prec", ["f64", "f32"])
@pytest.mark.parametrize("name", ["mixed2", "fit8x2_jit", "ex1dof_short", "ex1dof_interp", "fit8_e16"])
def test_golden_with_every_component_offered_to_the_split(name, prec):
    """large_threshold = 0 offers every component to the lane-block split; the cost model keeps it or hands the component
    to the tile kernels -- either way the reference's golden vectors must be reproduced."""
    z = golden(name, prec)
    case = wl.golden_case(name)
    out = run_case(capi.engine_class(prec), case, options={"large_threshold": "0"})
    p = out["_p"]
    assert p._query("lane_classes") + p._query("large_classes") >= 1
    E = len(case["excitations"])
    assert close(out["amplitudes"], z["amplitudes"], z["amplitudes"].size // E, prec, TOL_FWD)
    if "trajectory" in z.files:
        assert rel_err(out["trajectory"][1:], z["trajectory"][1:]) < TOL_FWD[prec]
    if "gradient" in z.files:
        assert close(out["gradient"], z["gradient"], z["gradient"].size // E, prec, TOL_ADJ)


def lattice_case(nx, ny, E, L):
    s = wl.lattice_system(nx, ny)
    return dict(systems=[s], jacobians=[wl.chain_jacobian(s, 3)], excitations=wl.random_sine_excitations(E, L, 16000, seed=9),
                sample_rate=16000, numsteps=0, interp=None, traj_save_steps=max(L // 5, 1))


@pytest.mark.parametrize("prec", ["f64", "f32"])
@pytest.mark.parametrize("shape", [(8, 8), (6, 10), (16, 16)])
def test_membrane_through_lane_groups_matches_oracle(shape, prec):
    """2-D lattice of oscillators: couplings reach +-1 and +-ny sites, so a lane exchanges a whole row of halo operands"""
    case = lattice_case(*shape, E=7, L=300)
    a = run_case(capi.engine_class(prec), case)
    b = run_case(_oracle(prec, "reference"), case)
    p = a["_p"]
    assert p._query("lane_classes") >= 2 and p._query("lane_halo_fwd") >= 6
    n, na = case["systems"][0]["n"], case["jacobians"][0]["n"]
    assert close(a["amplitudes"], b["amplitudes"], n, prec, TOL_FWD)
    assert rel_err(a["trajectory"], b["trajectory"]) < TOL_FWD[prec]
    assert close(a["gradient"], b["gradient"], na, prec, TOL_ADJ)


@pytest.mark.parametrize("prec", ["f64", "f32"])
def test_full_size_chain_properties(prec):
    """BASELINE config 4 at its full component size (256-DOF chain, 514 states, 518 adjoint states) through size-independent
    properties: duplicated excitations give bit-identical rows, the split and the tile kernels agree, the excitation-summed
    gradient equals the sum of the rows, and a second call reproduces the first bit for bit."""
    s = wl.chain_system(256)
    exc = wl.random_sine_excitations(256, 240, 16000, seed=11)
    case = dict(systems=[s], jacobians=[wl.chain_jacobian(s, 4)], excitations=np.concatenate([exc, exc]), sample_rate=16000,
                numsteps=0, interp=None, traj_save_steps=None)
    a = run_case(capi.engine_class(prec), case)
    p = a["_p"]
    # fp64: 64 lanes x 10 slots; fp32 (even E: pair kernels, two trajectories per thread): sized with every slot counted twice -> 64 lanes
    assert p._query("lane_classes") >= 2 and p._query("lane_G") == 64
    n, na, E = 514, 518, 512
    amp, grad = a["amplitudes"].reshape(E, n), a["gradient"].reshape(E, na)
    assert np.array_equal(amp[:256], amp[256:]) and np.array_equal(grad[:256], grad[256:])
    assert np.array_equal(a["amplitudes"], a["amplitudes_2nd"])
    p._setOption("call_order", "independent")      # (the reference carries state from call to call, quirk A7)
    assert np.array_equal(np.asarray(p.getGradient()), a["gradient"])
    assert np.all(np.isfinite(amp)) and np.all(np.isfinite(grad)) and np.max(np.abs(grad)) > 0
    red = np.asarray(p._getGradientReduced(0))
    assert rel_err(red, grad.sum(axis=0)) < (1e-12 if prec == "f64" else 2e-5)
    t = run_case(capi.engine_class(prec), case, options={"lanes": "off"})
    assert t["_p"]._query("lane_classes") == 0
    assert close(a["amplitudes"], t["amplitudes"], n, prec, TOL_FWD)
    assert close(a["gradient"], t["gradient"], na, prec, TOL_ADJ)


def test_gradient_through_lane_groups_matches_finite_differences():
    """exact (clean) schedule: d p1(T) / d Lambda_k from the adjoint sweep of a 60-DOF chain against a central finite
    difference of the forward probe -- the gradient the adjoint builder promises (pasif_b200/adjoint.py).  The adjoint ODE is
    integrated with the reference's scheme (forward state averaged at the half stages), which is not the algebraic
    transpose of the forward RK4 map: agreement is to discretisation error (4e-5 with the C oracle), not round-off."""
    s = wl.chain_system(60)
    exc = wl.random_sine_excitations(3, 400, 16000, seed=13)
    base = dict(excitations=exc, sample_rate=16000, numsteps=0, interp=None, traj_save_steps=None)
    a = run_case(capi.engine_class("f64", "clean"), dict(base, systems=[s], jacobians=[wl.chain_jacobian(s, 2)]))
    assert a["_p"]._query("lane_classes") >= 2
    n = s["n"]
    g = a["gradient"].reshape(3, n + 2)[:, n:]                    # two parameters: Lambda'_0, Lambda'_1 (pre-scaled)
    M = np.asarray(s["M"], dtype=np.float64)
    lidx = np.asarray(s["Lambda"][2]).reshape(-1, 4)
    for k in range(2):
        out_row = int(lidx[k, 0])
        vals = []
        for sign in (+1, -1):
            s2 = dict(s)
            lv = list(s["Lambda"][1])
            # pre-scaled coefficient Lambda' = -Lambda / M[out]: a step d in Lambda' is a step -d * M[out] in Lambda
            lam_p = -lv[k] / M[out_row]
            h = 1e-4 * abs(lam_p)
            lv[k] = -(lam_p + sign * h) * M[out_row]
            s2["Lambda"] = (s["Lambda"][0], lv, s["Lambda"][2])
            f = run_case(capi.engine_class("f64"), dict(base, systems=[s2], jacobians=None), want_grad=False)
            vals.append(f["amplitudes"].reshape(3, n)[:, n - 1])   # p1(T)
        fd = (vals[0] - vals[1]) / (2 * h)
        assert np.max(np.abs(fd - g[:, k])) < 1e-3 * max(np.max(np.abs(fd)), 1e-30), (k, fd, g[:, k])


@pytest.mark.parametrize("prec", ["f64", "f32"])
def test_mid_size_component_is_split_when_one_thread_would_spill(prec):
    """16-DOF chain (34 states, below large_threshold): in fp64 state + coefficients of the one-thread kernel need ~560
    registers, so the component is offered to the lane-block split (measured 16x on the gradient); in fp32 it still fits one
    thread.  Both must reproduce the oracle."""
    case = chain_case(16, False, E=9, L=500)
    a = run_case(capi.engine_class(prec), case)
    b = run_case(_oracle(prec, "reference"), case)
    assert a["_p"]._query("lane_classes") == (2 if prec == "f64" else 0)
    assert a["_p"]._query("large_classes") == 0
    n, na = case["systems"][0]["n"], case["jacobians"][0]["n"]
    assert close(a["amplitudes"], b["amplitudes"], n, prec, TOL_FWD)
    assert rel_err(a["trajectory"], b["trajectory"]) < TOL_FWD[prec]
    assert close(a["gradient"], b["gradient"], na, prec, TOL_ADJ)


@pytest.mark.parametrize("prec", ["f64", "f32"])
def test_lane_groups_with_interpolated_excitations(prec):
    """interpolation table (two force sub-samples per RK4 step, preprocessed half-step stream) through the lane groups"""
    case = chain_case(40, False, E=5, L=260)
    case["interp"] = (wl.linear_interpolation_table(8), 8)
    case["sample_rate"] = 4 * 16000
    case["numsteps"] = 4 * 250
    case["traj_save_steps"] = 100
    a = run_case(capi.engine_class(prec), case)
    b = run_case(_oracle(prec, "reference"), case)
    assert a["_p"]._query("lane_classes") >= 2
    n, na = case["systems"][0]["n"], case["jacobians"][0]["n"]
    assert close(a["amplitudes"], b["amplitudes"], n, prec, TOL_FWD)
    assert rel_err(a["trajectory"], b["trajectory"]) < TOL_FWD[prec]
    assert close(a["gradient"], b["gradient"], na, prec, TOL_ADJ)


@pytest.mark.parametrize("prec", ["f64", "f32"])
def test_lane_groups_hbm_sized_schedule_is_bit_identical(prec):
    """exact schedule: chunk trajectories of the lane classes kept from the forward sweep (tail_budget_gb) leave every bit
    of the gradient unchanged and save recompute launches; `save` stops the literal schedule after a chunk"""
    case = chain_case(40, False, E=6, L=900)
    ref = run_case(capi.engine_class(prec, "clean"), case, want_traj=False, options={"tail_budget_gb": "0"})
    assert ref["_p"]._query("lane_classes") >= 2 and ref["_p"]._query("stored_chunks") == 0
    out = run_case(capi.engine_class(prec, "clean"), case, want_traj=False, options={"tail_budget_gb": "1.0"})
    assert out["_p"]._query("stored_chunks") == out["_p"]._query("chunks") > 0
    assert np.array_equal(out["gradient"], ref["gradient"])
    assert out["_p"]._query("last_launches") < ref["_p"]._query("last_launches")
    lit = run_case(capi.engine_class(prec, "reference"), case, want_traj=False)
    part = np.asarray(lit["_p"].getGradient(20))                  # stop after chunk 20 of 30: a different, finite state
    assert part.shape == lit["gradient"].shape and np.all(np.isfinite(part)) and not np.array_equal(part, lit["gradient"])
    o = run_case(_oracle(prec, "reference"), case, want_traj=False)
    assert close(part, np.asarray(o["_p"].getGradient(20)), case["jacobians"][0]["n"], prec, TOL_ADJ)


@pytest.mark.parametrize("prec,G", [("f64", 256), ("f32", 128)])
def test_thousand_dof_chain_on_wide_groups(prec, G):
    """1000-DOF chain (2002 states -- beyond the shared-memory tile kernels' 1024 rows): in fp64 the group is wider than
    the default CTA, so the kernels are compiled for a CTA of 256 threads with one barrier per exchange"""
    case = chain_case(1000, False, E=3, L=200)
    a = run_case(capi.engine_class(prec), case)
    b = run_case(_oracle(prec, "reference"), case)
    assert a["_p"]._query("lane_classes") >= 2 and a["_p"]._query("lane_G") == G
    n, na = case["systems"][0]["n"], case["jacobians"][0]["n"]
    assert close(a["amplitudes"], b["amplitudes"], n, prec, TOL_FWD)
    assert rel_err(a["trajectory"], b["trajectory"]) < TOL_FWD[prec]
    assert close(a["gradient"], b["gradient"], na, prec, TOL_ADJ)


@pytest.mark.parametrize("prec", ["f64", "f32"])
def test_irregular_system_forced_through_lane_groups(prec):
    """a random banded system with long-range couplings, split with the cost model switched off: a poor split (halo operands
    at arbitrary lane offsets, mostly padding terms) that the kernels must still integrate exactly like the oracle"""
    from pasif_b200 import adjoint
    from test_lane_split import _random_system
    rng = np.random.default_rng(42)
    s = _random_system(rng, 60, band=2, extra=3)
    # keep the dynamics tame: damp the linear part
    kv, kr, kc = s["K"]
    s["K"] = ([v + (40.0 if r == c else 0.0) for v, r, c in zip(kv, kr, kc)], kr, kc)
    loss = np.zeros(60); loss[7] = 1.0
    j = adjoint.build_adjoint(s, loss, [("Lambda", 0), ("Gamma", 1), ("K", 0, 0)])
    case = dict(systems=[s], jacobians=[j], excitations=wl.random_sine_excitations(5, 300, 16000, seed=3), sample_rate=16000,
                numsteps=0, interp=None, traj_save_steps=50)
    a = run_case(capi.engine_class(prec), case, options={"lane_regs": "1000000", "lane_min_eff": "0", "lanes": "8"})
    b = run_case(_oracle(prec, "reference"), case)
    assert a["_p"]._query("lane_classes") >= 2 and a["_p"]._query("lane_halo_fwd") > 8
    assert close(a["amplitudes"], b["amplitudes"], 60, prec, TOL_FWD)
    assert rel_err(a["trajectory"], b["trajectory"]) < TOL_FWD[prec]
    assert close(a["gradient"], b["gradient"], 63, prec, TOL_ADJ)
