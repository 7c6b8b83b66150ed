"""CPU suite: the front end (pasif_b200/frontend.py) hands the native driver exactly what the reference's PASIf.py does."""
import json
import os

import numpy as np
import pytest

from conftest import ROOT, frontend_mod, load_reference_frontend
from pasif_b200 import frontend, workloads as wl
import pasif_oracle as po

FIXTURE = os.path.join(ROOT, "tests", "golden", "handoff_reference_frontend.json")
CASES = ["ex1dof", "fit8_e16", "fit8x2_jit", "mixed2"]


def _handoff(front_cls, mod, case, with_jac=True):
    exc = [np.array(r[:16]) for r in case["excitations"][:2]]
    p = front_cls(exc, case["sample_rate"], 0)
    p.setSystems(*wl.to_frontend_args(case["systems"], mod))
    if with_jac and case["jacobians"] is not None:
        p.setJacobian(*wl.to_frontend_args(case["jacobians"], mod, with_psi=True))
    return p.rec


def _same(a, b, path=""):
    if isinstance(a, (tuple, list)) and not np.isscalar(a) and len(a) and isinstance(a[0], (tuple, list, np.ndarray)):
        assert len(a) == len(b), path
        for i, (x, y) in enumerate(zip(a, b)):
            _same(x, y, f"{path}[{i}]")
    else:
        a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
        assert a.shape == b.shape, (path, a.shape, b.shape)
        assert np.array_equal(a, b), path


def _jsonable(rec):
    def conv(v):
        if isinstance(v, (tuple, list)):
            return [conv(x) for x in v]
        if isinstance(v, np.ndarray):
            return v.tolist()
        return v
    return {k: conv(v) for k, v in rec.items()}


@pytest.mark.parametrize("name", CASES)
def test_handoff_identical_to_reference_frontend(name):
    ref = load_reference_frontend(po.RecordingDriver)
    mine = frontend.make_frontend(po.RecordingDriver)
    case = wl.golden_case(name)
    a, b = _handoff(ref.PASIf, ref, case), _handoff(mine, frontend_mod, case)
    assert a.keys() == b.keys()
    for k in a:
        _same(a[k], b[k], k)


@pytest.mark.parametrize("name", CASES)
def test_handoff_matches_committed_fixture(name):
    """Same check where the reference is not mounted: fixture written by tests/golden/gen_handoff_fixture.py."""
    fix = json.load(open(FIXTURE))[name]
    rec = _jsonable(_handoff(frontend.make_frontend(po.RecordingDriver), frontend_mod, wl.golden_case(name)))
    assert rec.keys() == fix.keys()
    for k in fix:
        _same(rec[k], fix[k], k)


def test_probe_values_from_survey():
    """SURVEY.md 3.1 probe: M=diag(2,1,4,1) ... hands over K' data [0.5,1,-0.75,-1.25] etc."""
    from scipy.sparse import dia_matrix, coo_matrix
    P = frontend.make_frontend(po.RecordingDriver)
    p = P([np.ones(8)], 16000, 0)
    G = frontend.coo_tensor([4, 4, 4]); G.val = [-1.]; G.indices = [3, 0, 0]
    L = frontend.coo_tensor([4, 4, 4, 4]); L.val = [7.]; L.indices = [2, 0, 0, 0]
    p.setSystems([dia_matrix(([2., 1., 4., 1.], [0]), shape=(4, 4))],
                 [coo_matrix(([-1., -1.], ([0, 1], [2, 3])), shape=(4, 4))],
                 [coo_matrix(([3., 5.], ([2, 2], [0, 2])), shape=(4, 4))], [G], [L],
                 [np.array([0., 0., 1., 0.])], [np.zeros(4)])
    shape, data, indices, indptr = p.rec["fwdK"]
    assert list(data) == [0.5, 1.0, -0.75, -1.25] and list(indices) == [2, 3, 0, 2] and list(indptr) == [0, 1, 2, 4, 4]
    assert p.rec["fwdGamma"][1] == [1.0] and p.rec["fwdLambda"][1] == [-1.75]
    assert list(p.rec["fwdF"]) == [0., 0., 0.25, 0.]


def test_more_than_two_jacobians_and_no_input_mutation():
    """The reference front end raises for >= 3 Jacobians (PASIf.py:574) and overwrites vecForcePattern[0]; this one does neither."""
    rng = np.random.default_rng(3)
    systems = [wl.jitter(wl.fit8_system(), rng) for _ in range(4)]
    jacs = [wl.jitter(wl.fit8_jacobian(), rng) for _ in range(4)]
    args = wl.to_frontend_args(systems, frontend_mod)
    f0 = args[5][0].copy()
    p = frontend.make_frontend(po.RecordingDriver)([np.ones(8)], 16000, 0)
    p.setSystems(*args)
    assert np.array_equal(args[5][0], f0) and len(args[5][0]) == 8
    p.setJacobian(*wl.to_frontend_args(jacs, frontend_mod, with_psi=True))
    assert p.rec["bwdK"][0] == (48, 48) and p.rec["bwdPsi"][0] == [48, 48, 32, 32, 32]


def test_error_behaviour_matches_reference():
    P = frontend.make_frontend(po.RecordingDriver)
    p = P([np.ones(8)], 16000, 0)
    with pytest.raises(ValueError):
        p.getAmplitudes()
    with pytest.raises(ValueError):
        p.getTrajectory()
    with pytest.raises(ValueError):
        p.getGradient()
    with pytest.raises(ValueError):
        p.setJacobian(*wl.to_frontend_args([wl.fit8_jacobian()], frontend_mod, with_psi=True))
    with pytest.raises(ValueError):
        p.setInterpolationMatrix([np.ones(3)])
    with pytest.raises(ValueError):
        p.setModulationBuffer(np.zeros(0))
    with pytest.raises(AssertionError):
        args = wl.to_frontend_args([wl.fit8_system()], frontend_mod)
        args[5] = [np.zeros(3)]
        p.setSystems(*args)


def test_trajectory_repack_matches_reference_wrapper():
    ref = load_reference_frontend(po.RecordingDriver)
    mine = frontend.make_frontend(po.RecordingDriver)
    case = wl.golden_case("fit8x2_jit")
    exc = [np.array(r[:200]) for r in case["excitations"][:2]]
    a = ref.PASIf(exc, 16000, 0)
    a.setSystems(*wl.to_frontend_args(case["systems"], ref))
    b = mine(exc, 16000, 0)
    b.setSystems(*wl.to_frontend_args(case["systems"], frontend_mod))
    assert np.array_equal(a.getTrajectory(7), b.getTrajectory(7))
