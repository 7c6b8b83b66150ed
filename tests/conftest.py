import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box: pytest -m gpu)")


@pytest.fixture(scope="session", autouse=True)
def _native_build():
    """Build the in-tree native pieces once (no-op when up to date; nvcc cross-compiles without a GPU)."""
    import __graft_entry__ as g
    g.build()


def golden(name, prec):
    path = os.path.join(ROOT, "tests", "golden", f"ref_{name}_{prec}.npz")
    if not os.path.exists(path):
        pytest.skip(f"golden fixture {path} missing")
    return np.load(path)


def rel_err(a, b):
    """max |a - b| / max |b| over the whole output vector."""
    a = np.asarray(a, dtype=np.float64).reshape(-1)
    b = np.asarray(b, dtype=np.float64).reshape(-1)
    assert a.shape == b.shape, (a.shape, b.shape)
    if a.size == 0:
        return 0.0
    return float(np.max(np.abs(a - b)) / max(np.max(np.abs(b)), 1e-300))


def col_err(a, b, width):
    """per state variable: max over excitations of |a - b| / max over excitations of |b| (columns that are ~0 in the
    reference are normalised by the largest column instead)."""
    a = np.asarray(a, dtype=np.float64).reshape(-1, width)
    b = np.asarray(b, dtype=np.float64).reshape(-1, width)
    scale = np.max(np.abs(b), axis=0)
    scale = np.where(scale > 1e-12 * max(scale.max(), 1e-300), scale, max(scale.max(), 1e-300))
    return float(np.max(np.max(np.abs(a - b), axis=0) / scale))


class frontend_mod:
    """namespace handed to workloads.to_frontend_args for this package's front end"""
    from pasif_b200.frontend import coo_tensor


def run_case(driver_cls, case, want_traj=True, want_grad=True, options=None):
    """Drive a golden case through this package's front end on `driver_cls`; returns dict of outputs."""
    from pasif_b200 import frontend, workloads as wl
    P = frontend.make_frontend(driver_cls)
    p = P([np.array(r) for r in case["excitations"]], case["sample_rate"], case["numsteps"])
    for k, v in (options or {}).items():
        p._setOption(k, v)
    p.setSystems(*wl.to_frontend_args(case["systems"], frontend_mod))
    if case["interp"] is not None:
        p.setInterpolationMatrix([np.array(r) for r in case["interp"][0]])
    out = {"amplitudes": np.asarray(p.getAmplitudes()), "amplitudes_2nd": np.asarray(p.getAmplitudes())}
    if want_traj and case["traj_save_steps"]:
        out["trajectory"] = np.asarray(p.getTrajectory(case["traj_save_steps"]))
    if want_grad and case["jacobians"] is not None:
        p.setJacobian(*wl.to_frontend_args(case["jacobians"], frontend_mod, with_psi=True))
        out["gradient"] = np.asarray(p.getGradient())
    out["_p"] = p
    return out


def load_reference_frontend(driver_cls):
    """Import the UNMODIFIED reference front end (PASIf.py) on top of `driver_cls` posing as PASIfgpu.__GpuDriver.
    Source: /root/reference (build container) or the copy `make -C oracle ref` left in oracle/_ref (GPU box)."""
    import importlib.util
    import types
    for path in ("/root/reference/PASIf.py", os.path.join(ROOT, "oracle", "_ref", "PASIf.py")):
        if os.path.exists(path):
            break
    else:
        pytest.skip("reference front end not available")
    fake = types.ModuleType("PASIfgpu")
    setattr(fake, "__GpuDriver", driver_cls)
    saved = sys.modules.get("PASIfgpu")
    sys.modules["PASIfgpu"] = fake
    shim = os.path.join(ROOT, "oracle", "mpi4py_shim")
    if shim not in sys.path:
        sys.path.insert(0, shim)
    name = f"PASIf_reference_{id(driver_cls)}"
    spec = importlib.util.spec_from_file_location(name, path)
    m = importlib.util.module_from_spec(spec)
    sys.modules[name] = m
    try:
        spec.loader.exec_module(m)
    finally:
        if saved is not None:
            sys.modules["PASIfgpu"] = saved
        else:
            sys.modules.pop("PASIfgpu", None)
    return m
