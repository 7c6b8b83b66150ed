"""CPU suite: the C-ABI library loads, exports every symbol include/pasif_b200.h declares, fails loudly without a GPU,
and its host-side structural analysis (components -> classes -> compile-time traits) is right.  No compute calls."""
import ctypes
import os
import re
import subprocess
import sys

import numpy as np
import pytest

from conftest import ROOT, frontend_mod
from pasif_b200 import capi, frontend, workloads as wl


def _declared_symbols():
    text = open(os.path.join(ROOT, "include", "pasif_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(pasif_[a-z_0-9]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    names = _declared_symbols()
    assert len(names) >= 30
    L = ctypes.CDLL(capi.LIB_PATH)
    for n in names:
        assert hasattr(L, n), n
    assert set(capi.SYMBOLS) == set(names)


def test_library_has_sm100a_kernels():
    out = subprocess.run(["cuobjdump", "-lelf", capi.LIB_PATH], stdout=subprocess.PIPE, text=True).stdout
    assert "sm_100a" in out


def test_dense_operator_fast_path_kernels_fit_their_register_budget():
    """The update kernel of the dense-operator path is capped at 64 registers (4 resident CTAs per SM); its default
    instantiations (4 trajectories per thread, one instance per thread) must not pay for that with local-memory traffic:
    none in fp32, at most the stage-4 variant's few bytes in fp64.  The tcgen05 GEMM kernels never spill."""
    out = subprocess.run(["cuobjdump", "-res-usage", capi.LIB_PATH], stdout=subprocess.PIPE, text=True).stdout
    usage = {}
    name = None
    for line in out.splitlines():
        m = re.search(r"Function (\S+):", line)
        if m:
            name = m.group(1)
            continue
        m = re.search(r"REG:(\d+) STACK:(\d+)", line)
        if m and name:
            usage[name] = (int(m.group(1)), int(m.group(2)))
            name = None
    upd = {k: v for k, v in usage.items() if re.search(r"stage_update_kernelI[fd]Li4ELi\dELb[01]ELb1ELb[01]E", k)}
    assert len(upd) == 20, sorted(upd)                    # 2 types x (4 stages x {forward, adjoint} + 2 half stages)
    for k, (reg, stack) in upd.items():
        assert reg <= 64, (k, reg)
        assert stack == 0 or ("kernelId" in k and "Li4ELi4E" in k and stack <= 64), (k, stack)
    gemm = {k: v for k, v in usage.items() if "tcgen05" in k}
    assert len(gemm) >= 3
    for k, (reg, stack) in gemm.items():
        assert stack == 0, (k, stack)


def test_pybind_module_exposes_reference_interface():
    sys.path.insert(0, os.path.join(ROOT, "pasif_b200"))
    import PASIfgpu
    drv = getattr(PASIfgpu, "__GpuDriver")
    ref_methods = """_setFwdK _setFwdGamma _setFwdLambda _setFwdPsi _setFwdForcePattern _setFwdInitialConditions
    _allocateSystemOnDevice _setBwdK _setBwdGamma _setBwdLambda _setBwdPsi _setBwdForcePattern _setBwdInitialConditions
    _allocateAdjointOnDevice _loadExcitationsSet _setInterpolationMatrix _setModulationBuffer _getAmplitudes
    _getTrajectory _getGradient""".split()          # reference src/PASIfgpu.cpp:37-104
    for m in ref_methods:
        assert hasattr(drv, m), m


def _has_gpu():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


@pytest.mark.skipif(_has_gpu(), reason="only meaningful on a box without a GPU")
def test_no_cpu_fallback():
    with pytest.raises(capi.PasifError, match="no CPU fallback"):
        capi.Engine(np.ones((2, 10)), 16000, 10)
    sys.path.insert(0, os.path.join(ROOT, "pasif_b200"))
    import PASIfgpu
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        getattr(PASIfgpu, "__GpuDriver")([np.ones(10)], 16000, 10)


def _host(systems, jacobians=None, dtype="f64"):
    P = frontend.make_frontend(capi.host_engine_class(dtype))
    p = P([np.zeros(8)], 16000, 0)
    p.setSystems(*wl.to_frontend_args(systems, frontend_mod))
    if jacobians is not None:
        p.setJacobian(*wl.to_frontend_args(jacobians, frontend_mod, with_psi=True))
    return p


def test_missing_library_fails_loudly():
    """No native library -> ImportError naming the build step, in a fresh interpreter (nothing else is tried)."""
    code = ("import sys; sys.path.insert(0, %r)\n"
            "from pasif_b200 import capi\n"
            "import numpy as np\n"
            "capi.Engine(np.ones((2, 10)), 16000, 10)\n") % ROOT
    r = subprocess.run([sys.executable, "-c", code], env=dict(os.environ, PASIF_B200_LIB="/nonexistent/libpasif_b200.so"),
                       stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True)
    assert r.returncode != 0
    assert "ImportError" in r.stderr and "is not built" in r.stderr


def test_host_only_engine_refuses_to_solve():
    p = _host([wl.fit8_system()])
    with pytest.raises(capi.PasifError, match="host-only"):
        p.getAmplitudes()


def test_structural_classes():
    rng = np.random.default_rng(0)
    p = _host([wl.jitter(wl.fit8_system(), rng) for _ in range(5)], [wl.jitter(wl.fit8_jacobian(), rng) for _ in range(5)])
    assert (p._query("fwd_classes"), p._query("fwd_instances")) == (1, 5)
    assert (p._query("joint_classes"), p._query("joint_instances")) == (1, 5)
    assert p._query("n_fwd") == 40 and p._query("n_adj") == 60
    c = wl.golden_case("mixed2")
    p = _host(c["systems"], c["jacobians"])
    assert (p._query("fwd_classes"), p._query("fwd_instances")) == (2, 2)
    # ex1dof Jacobian: adjoint variable 3 is coupled to nothing -> its own (trivial) joint component
    p = _host([wl.ex1dof_system()], [wl.ex1dof_jacobian()])
    assert (p._query("joint_classes"), p._query("joint_instances")) == (2, 2)


def test_components_match_graph_connectivity():
    """Random block structure: the engine must find exactly the connected components of the coupling graph."""
    from scipy.sparse import coo_matrix
    from scipy.sparse.csgraph import connected_components
    rng = np.random.default_rng(5)
    n = 60
    eng = capi.host_engine_class("f64")(np.zeros((1, 8)), 16000, 8)
    rows = rng.integers(0, n, 40)
    cols = np.clip(rows + rng.integers(-2, 3, 40), 0, n - 1)
    K = coo_matrix((rng.normal(size=40), (rows, cols)), shape=(n, n)).tocsr()
    K.sort_indices()
    g_idx = np.stack([rng.integers(0, n, 10)] * 3, axis=1)
    g_idx[:, 2] = np.clip(g_idx[:, 0] + rng.integers(-1, 2, 10), 0, n - 1)
    eng._setFwdK((n, n), K.data, K.indices, K.indptr)
    eng._setFwdGamma([n] * 3, rng.normal(size=10), g_idx.reshape(-1))
    eng._setFwdLambda([n] * 4, [], [])
    eng._setFwdPsi([1] * 5, [], [])
    eng._setFwdForcePattern(np.zeros(n))
    eng._setFwdInitialConditions(np.zeros(n))
    eng._allocateSystemOnDevice()
    adj = coo_matrix((np.ones(40 + 20), (np.concatenate([rows, g_idx[:, 0], g_idx[:, 0]]),
                                         np.concatenate([cols, g_idx[:, 1], g_idx[:, 2]]))), shape=(n, n))
    ncomp, _ = connected_components(adj, directed=False)
    assert eng._query("fwd_instances") == ncomp


def test_prebuilt_pattern_table_is_current():
    """patterns_aot.inc is generated by the library itself; regenerating must reproduce the committed file."""
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "gen_aot_patterns.py")], stdout=subprocess.PIPE,
                       stderr=subprocess.STDOUT, text=True)
    assert r.returncode == 0, r.stdout
    assert "up to date" in r.stdout, r.stdout


def test_emitted_traits_compile_standalone(tmp_path):
    """The traits text the engine would hand to NVRTC compiles for sm_100a together with the kernel header (nvcc here)."""
    c = wl.golden_case("mixed2")
    p = _host(c["systems"], c["jacobians"])
    traits = p.emit_patterns()
    src = tmp_path / "jit_check.cu"
    body = ['#include "%s"' % os.path.join(ROOT, "pasif_b200", "csrc", "rk4_kernels.cuh"),
            "#define PASIF_AOT_FWD(TR, H) extern \"C\" __global__ void k_##TR(const pasif::FwdArgs<double> a) "
            "{ pasif::forward_body<double, TR, pasif::NoDense>(a); }",
            "#define PASIF_AOT_JOINT(TR, H) extern \"C\" __global__ void g_##TR(const pasif::FwdArgs<float> a) "
            "{ pasif::forward_body<float, TR::Fwd, TR::Adj>(a); } extern \"C\" __global__ void b_##TR(const "
            "pasif::BwdArgs<float> a) { pasif::backward_body<float, TR::Adj>(a); }", traits]
    src.write_text("\n".join(body))
    r = subprocess.run(["nvcc", "-gencode", "arch=compute_100a,code=sm_100a", "-std=c++17", "-cubin", "-o",
                        str(tmp_path / "o.cubin"), str(src), "-Xptxas", "-v"], stdout=subprocess.PIPE,
                       stderr=subprocess.STDOUT, text=True)
    assert r.returncode == 0, r.stdout
    assert "bytes spill" in r.stdout and not re.search(r"[1-9]\d* bytes spill", r.stdout), r.stdout


def test_embedded_kernel_header_is_in_sync():
    """rk4_kernels_text.inc (compiled into the library for NVRTC) is the current rk4_kernels.cuh."""
    import os
    from pasif_b200 import build as b
    text = open(os.path.join(b.CSRC, "rk4_kernels.cuh")).read()
    inc = open(os.path.join(b.CSRC, "rk4_kernels_text.inc")).read()
    assert inc.split('R"PASIFHDR(', 1)[1].rsplit(')PASIFHDR"', 1)[0] == text
