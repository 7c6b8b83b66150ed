"""GPU suite: excitation preprocessing -- the linear interpolation table generated on the
device, single-precision staging of the excitation set (bit-identical for an fp32 engine, optional with a stated error
for an fp64 engine), and the half as many bytes it moves over the host link."""
import numpy as np
import pytest

from conftest import frontend_mod, rel_err, col_err
from pasif_b200 import capi, frontend, workloads as wl

pytestmark = pytest.mark.gpu


def _front(prec, exc, sr, numsteps, options=None):
    p = frontend.make_frontend(capi.engine_class(prec))([np.array(r) for r in exc], sr, numsteps)
    for k, v in (options or {}).items():
        p._setOption(k, v)
    if options:
        p._loadExcitationsSet(np.asarray(exc), sr)     # staging options apply to the next upload
    return p


@pytest.mark.parametrize("prec", ["f64", "f32"])
@pytest.mark.parametrize("W", [2, 8, 32])
def test_device_generated_linear_table_equals_host_table(prec, W):
    """setLinearInterpolation(W) == setInterpolationMatrix(linear table built on the host), bit for bit: amplitudes, a
    trajectory and the gradient of the interpolated sweep."""
    case = wl.golden_case("ex1dof_interp")
    jac = [wl.ex1dof_jacobian()]
    outs = []
    for mode in ("host", "device"):
        p = _front(prec, case["excitations"], case["sample_rate"], 4000)
        p.setSystems(*wl.to_frontend_args(case["systems"], frontend_mod))
        if mode == "host":
            p.setInterpolationMatrix([np.array(r) for r in wl.linear_interpolation_table(W)])
        else:
            p.setLinearInterpolation(W)
        p.setJacobian(*wl.to_frontend_args(jac, frontend_mod, with_psi=True))
        outs.append((np.asarray(p.getAmplitudes()), np.asarray(p.getTrajectory(50)), np.asarray(p.getGradient())))
    for a, b in zip(*outs):
        assert np.array_equal(a, b)
    assert np.max(np.abs(outs[0][0])) > 0
    with pytest.raises((ValueError, capi.PasifError)):
        _front(prec, case["excitations"], case["sample_rate"], 100).setLinearInterpolation(6)


def test_fp32_engine_stages_single_precision_bit_identically():
    """An fp32 engine narrows the samples on the host while staging them (half the bytes over the link); the device used to
    do the same cast after the upload, so nothing changes in the results: golden parity of the fp32 fixtures covers it, here
    the byte count and a direct comparison with samples that are exactly representable in float."""
    exc = wl.random_sine_excitations(6, 3000, 16000, seed=3)
    exc32 = exc.astype(np.float32).astype(np.float64)              # already float-representable: narrowing is the identity
    s, j = [wl.fit8_system()], [wl.fit8_jacobian()]
    res = []
    for e in (exc, exc32):
        p = _front("f32", e, 16000, 0)
        p.setSystems(*wl.to_frontend_args(s, frontend_mod))
        p.setJacobian(*wl.to_frontend_args(j, frontend_mod, with_psi=True))
        res.append((np.asarray(p.getAmplitudes()), np.asarray(p.getGradient())))
        assert p._query("last_h2d_bytes") == exc.size * 4
    assert np.array_equal(res[0][0], res[1][0]) and np.array_equal(res[0][1], res[1][1])


def test_fp64_engine_optional_single_precision_excitations():
    """exc_storage=f32 on an fp64 engine: half the upload; the result equals an fp64 run on the float-rounded samples bit
    for bit, and differs from the full-precision run by the rounding of the samples only (<= 1e-6 relative on every state
    variable for the C2 system over 20 000 steps; each sample carries <= 6e-8)."""
    exc = wl.random_sine_excitations(16, 20000, 16000, seed=0)
    s, j = [wl.fit8_system()], [wl.fit8_jacobian()]
    out = {}
    for key, e, opt in (("full", exc, {}), ("narrow", exc, {"exc_storage": "f32"}),
                        ("rounded", exc.astype(np.float32).astype(np.float64), {})):
        p = _front("f64", e, 16000, 0, opt)
        p.setSystems(*wl.to_frontend_args(s, frontend_mod))
        p.setJacobian(*wl.to_frontend_args(j, frontend_mod, with_psi=True))
        out[key] = (np.asarray(p.getAmplitudes()), np.asarray(p.getGradient()), p._query("last_h2d_bytes"))
    assert out["full"][2] == exc.size * 8 and out["narrow"][2] == exc.size * 4
    assert np.array_equal(out["narrow"][0], out["rounded"][0]) and np.array_equal(out["narrow"][1], out["rounded"][1])
    assert 0 < col_err(out["narrow"][0], out["full"][0], 8) < 1e-6
    assert 0 < col_err(out["narrow"][1], out["full"][1], 12) < 1e-6
    assert rel_err(out["narrow"][1], out["full"][1]) < 1e-7
