"""GPU suite (pytest -m gpu, on the B200 box): the CUDA path, called through the C ABI, against
  * golden vectors produced by the compiled unmodified reference on a B200 (tests/golden/ref_*.npz),
  * the C oracle on the same seeded inputs,
  * size-independent properties at larger sizes (replica equality, linearity, reduction consistency, idempotence).
Tolerances (north star: <= 1e-12 relative in double, <= 1e-5 relative in float):
  fp64  every state variable: max_e |a - b| <= 1e-12 * max_e |b|                     (col_err, per variable)
  fp32  whole output vector:  max |a - b|   <= 1e-5 * max |b|, forward AND adjoint    (rel_err)
        and per variable <= 5e-4.  Measured (profiles/PARITY.md, every fixture of the unmodified reference): the CUDA path
        is within 3.6e-6 (forward) / 7.7e-10 (adjoint) of the reference's fp32 output on the whole vector.  Per variable
        the worst case is 1.5e-4 (fit8x2_jit amplitudes: a variable that ends near a zero crossing of its oscillation
        carries the fp32 round-off of 20 000 steps relative to a small final value) -- and the literal C restatement of
        the reference, which differs from the reference's GPU build only in rounding order, is exactly as far from it
        (1.5e-4), while the reference's own fp32 build is up to 1.7e-3 per variable from its fp64 build.  So 1e-5 per
        VARIABLE is below what two fp32 builds of the same arithmetic can agree on; it holds on the whole vector."""
import os
import sys

import numpy as np
import pytest

from conftest import ROOT, golden, rel_err, col_err, run_case, frontend_mod, load_reference_frontend
from pasif_b200 import capi, frontend, workloads as wl
from pasif_oracle_c import CRecordingDriver

pytestmark = pytest.mark.gpu

TOL_FWD = {"f64": 1e-12, "f32": 1e-5}
TOL_ADJ = {"f64": 1e-12, "f32": 1e-5}
TOL_COL32 = 5e-4


def close(a, b, width, prec, tol):
    """fp64: per-variable relative error; fp32: whole-vector relative error + a looser per-variable bound."""
    if prec == "f64":
        return col_err(a, b, width) < tol[prec]
    return rel_err(a, b) < tol[prec] and col_err(a, b, width) < TOL_COL32


def _oracle_driver(prec, compat="reference"):
    class D(CRecordingDriver):
        dtype = np.float64 if prec == "f64" else np.float32
    D.compat = compat
    return D


@pytest.mark.parametrize("prec", ["f64", "f32"])
@pytest.mark.parametrize("name", wl.GOLDEN_CASES)
def test_matches_reference_golden(name, prec):
    z = golden(name, prec)
    case = wl.golden_case(name)
    out = run_case(capi.engine_class(prec), case)
    E = len(case["excitations"])
    nf = z["amplitudes"].size // E
    assert close(out["amplitudes"], z["amplitudes"], nf, prec, TOL_FWD)
    assert np.array_equal(out["amplitudes"], out["amplitudes_2nd"])          # state reset between calls
    if "trajectory" in z.files:
        assert out["trajectory"].shape == z["trajectory"].shape
        assert rel_err(out["trajectory"][1:], z["trajectory"][1:]) < TOL_FWD[prec]
        assert np.allclose(out["trajectory"][0], z["trajectory"][0], rtol=1e-15)
    if "gradient" in z.files:
        assert close(out["gradient"], z["gradient"], z["gradient"].size // E, prec, TOL_ADJ)


@pytest.mark.parametrize("name", ["fit8_e16", "fit8x2_jit", "mixed2", "duffing8"])
def test_fp32_deviation_is_rounding_noise(name):
    """The CUDA fp32 path is no further from the reference's fp32 GPU output than the literal C restatement is
    (within a factor 4): what is left is accumulated fp32 round-off, not an algorithmic difference."""
    z = golden(name, "f32")
    case = wl.golden_case(name)
    gpu = run_case(capi.engine_class("f32"), case, want_traj=False)
    cpu = run_case(_oracle_driver("f32"), case, want_traj=False)
    for key in ("amplitudes", "gradient"):
        if key in z.files:
            assert rel_err(gpu[key], z[key]) <= 4 * rel_err(cpu[key], z[key]) + 2e-6, key


@pytest.mark.parametrize("prec", ["f64", "f32"])
@pytest.mark.parametrize("compat", ["reference", "clean"])
@pytest.mark.parametrize("name", ["fit8x2_jit", "mixed2", "ex1dof_short", "ex1dof_sq", "ex1dof_interp"])
def test_matches_oracle(name, compat, prec):
    case = wl.golden_case(name)
    a = run_case(capi.engine_class(prec, compat), case)
    b = run_case(_oracle_driver(prec, compat), case)
    E = len(case["excitations"])
    assert close(a["amplitudes"], b["amplitudes"], a["amplitudes"].size // E, prec, TOL_FWD)
    if "trajectory" in a:
        assert rel_err(a["trajectory"], b["trajectory"]) < TOL_FWD[prec]
    if "gradient" in a:
        assert close(a["gradient"], b["gradient"], a["gradient"].size // E, prec, TOL_ADJ)


@pytest.mark.parametrize("prec", ["f64", "f32"])
def test_call_order_matches_reference(prec):
    """Repeated solver calls on ONE driver, pinned by a fixture of the unmodified reference
    (tests/golden/ref_calls_ex1dof_short_*.npz): a second _getGradient without setJacobian continues from the adjoint state
    and the forward state the first one left behind (g2 != g1), getAmplitudes after a gradient starts from that forward
    state (a1 != a0) and resets it afterwards (a2 == a0)."""
    z = golden("calls_ex1dof_short", prec)
    case = wl.golden_case("calls_ex1dof_short")
    p = frontend.make_frontend(capi.engine_class(prec))([np.array(r) for r in case["excitations"]], case["sample_rate"], case["numsteps"])
    p.setSystems(*wl.to_frontend_args(case["systems"], frontend_mod))
    out = wl.run_call_sequence(p, case, frontend_mod)
    assert not np.array_equal(z["g1"], z["g2"]) and not np.array_equal(z["a0"], z["a1"]) and np.array_equal(z["a0"], z["a2"])
    for k in ("a0", "a1", "a2"):
        assert close(out[k], z[k], 4, prec, TOL_FWD), k
    for k in ("g1", "g2"):
        assert close(out[k], z[k], 6, prec, TOL_ADJ), k
    assert np.array_equal(out["a0"], out["a2"])
    # option call_order=independent: every call starts from the initial conditions
    q = frontend.make_frontend(capi.engine_class(prec))([np.array(r) for r in case["excitations"]], case["sample_rate"], case["numsteps"])
    q._setOption("call_order", "independent")
    q.setSystems(*wl.to_frontend_args(case["systems"], frontend_mod))
    ind = wl.run_call_sequence(q, case, frontend_mod)
    assert np.array_equal(ind["g1"], ind["g2"]) and np.array_equal(ind["a0"], ind["a1"]) and np.array_equal(ind["g1"], out["g1"])


@pytest.mark.parametrize("prec", ["f64", "f32"])
def test_call_order_reseeding_matches_oracle(prec):
    """The shipped example's re-seeding loop (examples/FullInterfaceTesting_1dof.py:233-237: _getGradient(save), then
    setJacobian with the returned state as the adjoint initial condition).  The unmodified reference dies in this sequence
    on a B200 (illegal memory access after the second setJacobian: its once-captured backward graph keeps pointers to the
    freed adjoint buffers, quirk A9), so the pin is the literal restatement, whose call-order state matches the
    reference's fixture in test_oracle.py."""
    case = wl.golden_case("calls_ex1dof_e1")
    outs = []
    for drv in (capi.engine_class(prec), _oracle_driver(prec)):
        p = frontend.make_frontend(drv)([np.array(r) for r in case["excitations"]], case["sample_rate"], case["numsteps"])
        p.setSystems(*wl.to_frontend_args(case["systems"], frontend_mod))
        outs.append(wl.run_call_sequence(p, case, frontend_mod))
    a, b = outs
    assert set(a) == set(b) and "s2" in a
    for k in a:
        assert close(a[k], b[k], 4 if k[0] == "a" else 6, prec, TOL_FWD if k[0] == "a" else TOL_ADJ), k


def test_save_argument_stops_after_chunk():
    """_getGradient(save) breaks out of the chunk loop after chunk `save` (reference src/__GpuDriver.cu:571-573)."""
    case = wl.golden_case("ex1dof_short")
    a = run_case(capi.engine_class("f64"), case, want_traj=False)["_p"]
    b = run_case(_oracle_driver("f64"), case, want_traj=False)["_p"]
    for save in (69, 40, 1):
        assert col_err(a._getGradient(save), b._getGradient(save), 6) < 1e-12


@pytest.mark.parametrize("prec", ["f64", "f32"])
def test_hbm_sized_schedule_is_bit_identical(prec):
    """Exact schedule (compat=clean): chunk trajectories kept from the forward sweep in free device memory
    (option tail_budget_gb: auto / a byte budget / 0 = always recompute) never change a bit of the result; only the
    number of recompute launches changes."""
    case = wl.golden_case("fit8x2_jit")
    E = len(case["excitations"])
    n = case["numsteps"] or len(case["excitations"][0])
    chunk = int(np.ceil(np.sqrt(n)))
    nsp = (n + chunk - 1) // chunk
    chunk_gb = chunk * 4 * (2 * E) * (8 if prec == "f64" else 4) / 1e9           # NV = 4 needed forward variables
    ref = run_case(capi.engine_class(prec, "clean"), case, want_traj=False, options={"tail_budget_gb": "0"})
    assert ref["_p"]._query("stored_chunks") == 0 and ref["_p"]._query("chunks") == nsp
    launches_ref = ref["_p"]._query("last_launches")
    for budget, stored in ((repr(1e3 * chunk_gb * nsp), nsp), (repr(3.5 * chunk_gb), 3)):
        out = run_case(capi.engine_class(prec, "clean"), case, want_traj=False, options={"tail_budget_gb": budget})
        assert out["_p"]._query("stored_chunks") == stored
        assert np.array_equal(out["gradient"], ref["gradient"])
        assert out["_p"]._query("last_launches") < launches_ref
    # automatic sizing leaves a batch too small to saturate the GPU on the (L2-resident) recompute path
    auto = run_case(capi.engine_class(prec, "clean"), case, want_traj=False, options={"tail_budget_gb": "auto"})
    assert auto["_p"]._query("stored_chunks") == 0 and np.array_equal(auto["gradient"], ref["gradient"])
    # the literal reference schedule replays a shifted trajectory (quirk A1): nothing can be kept from the sweep
    lit = run_case(capi.engine_class(prec, "reference"), case, want_traj=False)
    assert lit["_p"]._query("stored_chunks") == 0


def test_runtime_specialised_pattern_nvrtc():
    """A sparsity pattern with no prebuilt kernel goes through NVRTC (never a CPU or interpreted fallback)."""
    case = wl.golden_case("mixed2")
    out = run_case(capi.engine_class("f64"), case)
    p = out["_p"]
    assert p._query("jit_compiles") + p._query("jit_cache_hits") >= 1
    ref = run_case(_oracle_driver("f64"), case)
    assert col_err(out["gradient"], ref["gradient"], 12) < 1e-12
    # with jit off the same system must be refused loudly
    P = frontend.make_frontend(capi.engine_class("f64"))
    q = P([np.array(r) for r in case["excitations"]], 16000, 0)
    q._setOption("jit", "off")
    with pytest.raises(capi.PasifError, match="no prebuilt kernel"):
        q.setSystems(*wl.to_frontend_args(case["systems"], frontend_mod))


def test_unmodified_reference_frontend_on_pybind_module():
    """Drop-in: the reference's own PASIf.py drives this engine through the pybind11 class PASIfgpu.__GpuDriver."""
    sys.path.insert(0, os.path.join(ROOT, "pasif_b200"))
    import PASIfgpu
    ref = load_reference_frontend(getattr(PASIfgpu, "__GpuDriver"))
    case = wl.golden_case("fit8_e16")
    z = golden("fit8_e16", "f64")
    p = ref.PASIf([np.array(r) for r in case["excitations"]], case["sample_rate"], 0)
    p.setSystems(*wl.to_frontend_args(case["systems"], ref))
    amp = p.getAmplitudes()
    assert isinstance(amp, list)
    assert col_err(amp, z["amplitudes"], 8) < 1e-12
    traj = p.getTrajectory(case["traj_save_steps"])
    assert rel_err(traj[1:], z["trajectory"][1:]) < 1e-12
    p.setJacobian(*wl.to_frontend_args(case["jacobians"], ref, with_psi=True))
    assert col_err(p.getGradient(), z["gradient"], 12) < 1e-12


def test_full_size_properties_many_systems():
    """Config-3-like shape (many jittered systems x 64 excitations, fp32): replicas of one system agree with a
    single-system run, a second call reproduces the first bit for bit, and the device-side excitation reduction
    equals the host sum."""
    rng = np.random.default_rng(1)
    base = [wl.jitter(wl.fit8_system(), rng) for _ in range(3)]
    jbase = [wl.jitter(wl.fit8_jacobian(), rng) for _ in range(3)]
    S = 300
    pick = rng.integers(0, 3, S)
    exc = wl.random_sine_excitations(64, 4000, 16000, seed=2)
    P = frontend.make_frontend(capi.engine_class("f32"))
    p = P([np.array(r) for r in exc], 16000, 0)
    p.setSystems(*wl.to_frontend_args([base[i] for i in pick], frontend_mod))
    p.setJacobian(*wl.to_frontend_args([jbase[i] for i in pick], frontend_mod, with_psi=True))
    p._setOption("call_order", "independent")      # every call starts from the initial conditions (reference: quirk A7)
    amp = np.asarray(p.getAmplitudes()).reshape(64, S, 8)
    grad = np.asarray(p.getGradient()).reshape(64, S, 12)
    assert np.array_equal(np.asarray(p.getGradient()).reshape(64, S, 12), grad)
    for k in range(3):
        q = P([np.array(r) for r in exc], 16000, 0)
        q._setOption("call_order", "independent")
        q.setSystems(*wl.to_frontend_args([base[k]], frontend_mod))
        q.setJacobian(*wl.to_frontend_args([jbase[k]], frontend_mod, with_psi=True))
        a1 = np.asarray(q.getAmplitudes()).reshape(64, 8)
        g1 = np.asarray(q.getGradient()).reshape(64, 12)
        idx = np.nonzero(pick == k)[0]
        assert np.array_equal(amp[:, idx, :], np.broadcast_to(a1[:, None, :], (64, len(idx), 8)))
        assert np.array_equal(grad[:, idx, :], np.broadcast_to(g1[:, None, :], (64, len(idx), 12)))
    red = np.asarray(p._getGradientReduced(0))
    host = grad.astype(np.float32).sum(axis=0, dtype=np.float64).reshape(-1)
    assert rel_err(red, host) < 1e-5


def test_linearity_of_a_linear_system():
    """Gamma = Lambda = 0: amplitudes are linear in the excitation (checked at full length, 78 001 steps)."""
    s = wl.ex1dof_system(k=400.0, b=3.0)
    s["Gamma"] = (s["Gamma"][0], [], [])
    s["Lambda"] = (s["Lambda"][0], [], [])
    e = wl.random_sine_excitations(2, 78001, 16000, seed=4)
    exc = np.stack([e[0], e[1], 2.0 * e[0] - 3.0 * e[1]])
    P = frontend.make_frontend(capi.engine_class("f64"))
    p = P([np.array(r) for r in exc], 16000, 0)
    p.setSystems(*wl.to_frontend_args([s], frontend_mod))
    a = np.asarray(p.getAmplitudes()).reshape(3, 4)
    assert np.max(np.abs(a[2] - (2.0 * a[0] - 3.0 * a[1]))) < 1e-10 * np.max(np.abs(a))


def test_edge_cases():
    P = frontend.make_frontend(capi.engine_class("f64"))
    # numsteps smaller than 4 -> sqrt schedule has a single setpoint: the reference returns the adjoint IC untouched
    case = wl.golden_case("ex1dof_short")
    for n in (1, 3):
        case["numsteps"] = n
        a = run_case(capi.engine_class("f64"), case, want_traj=False)
        b = run_case(_oracle_driver("f64"), case, want_traj=False)
        assert col_err(a["amplitudes"], b["amplitudes"], 4) < 1e-12
        assert np.array_equal(a["gradient"], b["gradient"])
    # numsteps larger than the excitation length: the force gate switches the drive off past L (src/kernels.cu:84)
    case = wl.golden_case("ex1dof_short")
    case["numsteps"] = 7000
    a = run_case(capi.engine_class("f64"), case, want_grad=False)
    b = run_case(_oracle_driver("f64"), case, want_grad=False)
    assert col_err(a["amplitudes"], b["amplitudes"], 4) < 1e-12
    # a state variable nothing couples to, force entries below eps are dropped, E = 1, saveSteps not dividing numsteps
    s = wl.ex1dof_system(k=50.0, b=1.0)
    s["F"] = np.array([0., 1e-17, 1., 0.])
    iso = dict(systems=[s], jacobians=None, excitations=wl.sine_excitations([7.], 1000, 16000), sample_rate=16000,
               numsteps=0, interp=None, traj_save_steps=7)
    a = run_case(capi.engine_class("f64"), iso)
    b = run_case(_oracle_driver("f64"), iso)
    assert col_err(a["amplitudes"], b["amplitudes"], 4) < 1e-12
    assert a["trajectory"].shape == b["trajectory"].shape and rel_err(a["trajectory"], b["trajectory"]) < 1e-12
    # ragged excitations are rejected
    with pytest.raises((ValueError, capi.PasifError)):
        P([np.ones(5), np.ones(6)], 16000, 0)


def test_fma_peak_probe_runs():
    f64, f32 = capi.fma_peak(0, "f64"), capi.fma_peak(0, "f32")
    assert 5 < f64 < 80 and 20 < f32 < 160
