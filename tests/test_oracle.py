"""CPU suite: the oracle against the reference's known-answer cases and against golden vectors produced by the compiled,
unmodified reference on a B200 (tests/golden/ref_*.npz, generator tests/golden/gen_reference_golden.py)."""
import numpy as np
import pytest

from conftest import golden, rel_err, col_err, run_case
from pasif_b200 import workloads as wl
import pasif_oracle as po
from pasif_oracle_c import CRecordingDriver

# tolerance between two implementations of the same arithmetic with different rounding order
TOL = {"f64": 1e-12, "f32": 5e-5}


def test_reference_kernel_known_answers():
    """The only known-answer vectors in the reference repository: src/utest.cu:220-251 (updateSlope) and :255-296 (integrate)."""
    q = np.ones(3)
    rk = np.array([1., 2., 3.])
    assert np.array_equal(po.update_slope(q, rk, 2.0), [3., 5., 7.])
    assert np.array_equal(po.integrate(q, rk, rk, rk, rk, 2.0), [13., 25., 37.])


def _driver(base, prec):
    class D(base):
        dtype = np.float64 if prec == "f64" else np.float32
    return D


@pytest.mark.parametrize("prec", ["f64", "f32"])
@pytest.mark.parametrize("name", ["ex1dof_short", "ex1dof_sq", "ex1dof_interp"])
def test_numpy_oracle_matches_reference_golden(name, prec):
    z = golden(name, prec)
    out = run_case(_driver(po.RecordingDriver, prec), wl.golden_case(name))
    for key in ("amplitudes", "amplitudes_2nd", "trajectory", "gradient"):
        if key in z.files:
            assert rel_err(out[key], z[key]) < TOL[prec], key


@pytest.mark.parametrize("prec", ["f64", "f32"])
@pytest.mark.parametrize("name", [c for c in wl.GOLDEN_CASES if c != "fit8_c2"])
def test_c_oracle_matches_reference_golden(name, prec):
    z = golden(name, prec)
    case = wl.golden_case(name)
    out = run_case(_driver(CRecordingDriver, prec), case)
    E = len(case["excitations"])
    for key in ("amplitudes", "amplitudes_2nd", "trajectory"):
        if key in z.files:
            assert rel_err(out[key], z[key]) < TOL[prec], key
    if "gradient" in z.files:
        assert col_err(out["gradient"], z["gradient"], z["gradient"].size // E) < TOL[prec]


def test_c_oracle_matches_reference_golden_config2_subset():
    """BASELINE config 2 (E = 1024): first 64 excitations only, to keep the CPU suite short."""
    z = golden("fit8_c2", "f64")
    case = wl.golden_case("fit8_c2")
    case["excitations"] = case["excitations"][:64]
    out = run_case(_driver(CRecordingDriver, "f64"), case)
    assert rel_err(out["amplitudes"], z["amplitudes"][:64 * 8]) < 1e-12
    assert col_err(out["gradient"], z["gradient"][:64 * 12], 12) < 1e-12


@pytest.mark.parametrize("prec", ["f64", "f32"])
@pytest.mark.parametrize("base", [po.RecordingDriver, CRecordingDriver])
def test_oracle_call_order_matches_reference(base, prec):
    """Call-order state against a fixture of the unmodified reference: amplitudes, two gradients in a row,
    amplitudes twice more on one driver (pasif_b200.workloads.run_call_sequence)."""
    from conftest import frontend_mod
    from pasif_b200 import frontend
    z = golden("calls_ex1dof_short", prec)
    case = wl.golden_case("calls_ex1dof_short")
    p = frontend.make_frontend(_driver(base, prec))([np.array(r) for r in case["excitations"]], case["sample_rate"], case["numsteps"])
    p.setSystems(*wl.to_frontend_args(case["systems"], frontend_mod))
    out = wl.run_call_sequence(p, case, frontend_mod)
    assert not np.array_equal(z["g1"], z["g2"]) and not np.array_equal(z["a0"], z["a1"])
    for k in ("a0", "g1", "g2", "a1", "a2"):
        assert rel_err(out[k], z[k]) < TOL[prec], k
        assert col_err(out[k], z[k], 4 if k[0] == "a" else 6) < (1e-12 if prec == "f64" else 5e-5), k


def _random_case(seed, n_sys=2, L=400):
    rng = np.random.default_rng(seed)
    return dict(systems=[wl.jitter(wl.fit8_system(), rng, 0.3) for _ in range(n_sys)],
                jacobians=[wl.jitter(wl.fit8_jacobian(), rng, 0.3) for _ in range(n_sys)],
                excitations=wl.random_sine_excitations(3, L, 16000, seed=seed + 1), sample_rate=16000, numsteps=0,
                interp=None, traj_save_steps=37)


@pytest.mark.parametrize("compat", ["reference", "clean"])
def test_numpy_and_c_oracle_agree(compat):
    case = _random_case(11)

    class A(po.RecordingDriver):
        pass

    class B(CRecordingDriver):
        pass
    A.compat = B.compat = compat
    a, b = run_case(A, case), run_case(B, case)
    for key in ("amplitudes", "trajectory", "gradient"):
        assert rel_err(a[key], b[key]) < 1e-13, key


def test_clean_gradient_self_consistency():
    """With the schedule quirks removed the adjoint of the shipped 1-DOF example returns the forward probe integral
    (SURVEY.md Appendix A.3): gradient[4] == amplitudes[3] == int x^2 dt."""
    case = wl.golden_case("ex1dof")
    case["excitations"] = case["excitations"][:, :20001]

    class D(CRecordingDriver):
        compat = "clean"
    out = run_case(D, case, want_traj=False)
    assert abs(out["gradient"][4] - out["amplitudes"][3]) < 1e-9 * abs(out["amplitudes"][3])


def test_reference_quirk_magnitudes_are_visible():
    """The literal schedule differs measurably from the exact one (chunk 0 skipped etc.): parity must be literal."""
    case = wl.golden_case("ex1dof_short")

    class C(CRecordingDriver):
        compat = "clean"
    lit = run_case(CRecordingDriver, case, want_traj=False)["gradient"]
    cln = run_case(C, case, want_traj=False)["gradient"]
    assert col_err(lit, cln, 6) > 1e-3
