"""ctypes binding of the C ABI in ``include/pasif_b200.h`` (``libpasif_b200.so``).

``Engine`` mirrors the pybind class ``PASIfgpu.__GpuDriver`` method for method (same ``_underscore`` names, so the
front end in ``pasif_b200.frontend`` can sit on either), but moves data as NumPy host buffers without Python-list
conversion; the parity tests and ``bench.py`` call the engine through it.  There is no CPU fallback: constructing an
``Engine`` without a CUDA device raises, unless ``device=-1`` (host-only structural analysis).
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("PASIF_B200_LIB") or os.path.join(HERE, "libpasif_b200.so")   # override: A/B builds of the library

SYMBOLS = """pasif_create pasif_destroy pasif_last_error pasif_load_excitations
pasif_set_fwd_k pasif_set_fwd_gamma pasif_set_fwd_lambda pasif_set_fwd_psi pasif_set_fwd_force pasif_set_fwd_ic pasif_alloc_system
pasif_set_bwd_k pasif_set_bwd_gamma pasif_set_bwd_lambda pasif_set_bwd_psi pasif_set_bwd_force pasif_set_bwd_ic pasif_alloc_adjoint
pasif_set_interp pasif_set_interp_linear pasif_set_modulation pasif_set_option
pasif_get_amplitudes pasif_trajectory_size pasif_get_trajectory pasif_get_gradient pasif_get_gradient_reduced
pasif_get_gradient_reduced_device pasif_reduce_last_gradient_device pasif_query pasif_emit_patterns pasif_fma_peak pasif_debug_gemm pasif_debug_jit_log""".split()

_lib = None
_dp = C.POINTER(C.c_double)
_up = C.POINTER(C.c_uint32)


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError(f"{LIB_PATH} is not built; run `python -c 'import __graft_entry__ as g; g.build()'`")
        L = C.CDLL(LIB_PATH)
        L.pasif_create.argtypes = [C.POINTER(C.c_void_p), _dp, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32, C.c_int, C.c_int]
        L.pasif_destroy.argtypes = [C.c_void_p]
        L.pasif_destroy.restype = None
        L.pasif_last_error.argtypes = [C.c_void_p]
        L.pasif_last_error.restype = C.c_char_p
        L.pasif_load_excitations.argtypes = [C.c_void_p, _dp, C.c_uint32, C.c_uint32, C.c_uint32]
        for n in ("pasif_set_fwd_k", "pasif_set_bwd_k"):
            getattr(L, n).argtypes = [C.c_void_p, C.c_uint32, C.c_uint32, _dp, _up, _up, C.c_uint32]
        for n in ("gamma", "lambda", "psi"):
            for d in ("fwd", "bwd"):
                getattr(L, f"pasif_set_{d}_{n}").argtypes = [C.c_void_p, _up, _dp, _up, C.c_uint32]
        for n in ("pasif_set_fwd_force", "pasif_set_fwd_ic", "pasif_set_bwd_force", "pasif_set_bwd_ic", "pasif_set_modulation"):
            getattr(L, n).argtypes = [C.c_void_p, _dp, C.c_uint32]
        L.pasif_alloc_system.argtypes = [C.c_void_p]
        L.pasif_alloc_adjoint.argtypes = [C.c_void_p]
        L.pasif_set_interp.argtypes = [C.c_void_p, _dp, C.c_uint32, C.c_uint32]
        L.pasif_set_interp_linear.argtypes = [C.c_void_p, C.c_uint32]
        L.pasif_set_option.argtypes = [C.c_void_p, C.c_char_p, C.c_char_p]
        L.pasif_get_amplitudes.argtypes = [C.c_void_p, _dp, C.c_uint64]
        L.pasif_trajectory_size.argtypes = [C.c_void_p, C.c_uint32]
        L.pasif_trajectory_size.restype = C.c_uint64
        L.pasif_get_trajectory.argtypes = [C.c_void_p, C.c_uint32, _dp, C.c_uint64]
        L.pasif_get_gradient.argtypes = [C.c_void_p, C.c_uint32, _dp, C.c_uint64]
        L.pasif_get_gradient_reduced.argtypes = [C.c_void_p, C.c_uint32, _dp, C.c_uint64]
        L.pasif_get_gradient_reduced_device.argtypes = [C.c_void_p, C.c_uint32, C.c_void_p, C.c_uint64]
        L.pasif_reduce_last_gradient_device.argtypes = [C.c_void_p, C.c_void_p, C.c_uint64]
        L.pasif_query.argtypes = [C.c_void_p, C.c_char_p, _dp]
        L.pasif_emit_patterns.argtypes = [C.c_void_p, C.c_char_p, C.c_uint64, C.POINTER(C.c_uint64)]
        L.pasif_debug_jit_log.argtypes = [C.c_void_p, C.c_char_p, C.c_char_p, C.c_uint64, C.POINTER(C.c_uint64)]
        L.pasif_fma_peak.argtypes = [C.c_int, C.c_int, _dp]
        L.pasif_debug_gemm.argtypes = [C.c_int, C.c_int, C.c_int, C.c_uint32, C.c_uint64, _dp, _dp, _dp]
        _lib = L
    return _lib


def _d(a):
    a = np.ascontiguousarray(a, dtype=np.float64)
    return a, a.ctypes.data_as(_dp)


def _u(a):
    a = np.ascontiguousarray(a, dtype=np.uint32)
    return a, a.ctypes.data_as(_up)


class PasifError(RuntimeError):
    pass


class Engine:
    """Same constructor and `_methods` as ``PASIfgpu.__GpuDriver`` (reference src/PASIfgpu.cpp:26-104)."""

    def __init__(self, excitationSet_, sampleRate_, numsteps_, dCompute_=False, dSystem_=False, dSolver_=False,
                 GPUId_=0, dtype_=None, device=None):
        self._h = C.c_void_p()
        dtype_ = dtype_ or os.environ.get("PASIF_DTYPE", "f64")
        self.dtype = {"f64": 0, "f32": 1, "float64": 0, "float32": 1}[str(dtype_)]
        exc = np.ascontiguousarray(np.asarray(excitationSet_, dtype=np.float64))
        if exc.ndim != 2:
            raise ValueError("Excitations vectors are not of the same size")
        dev = int(GPUId_) if device is None else int(device)
        keep, p = _d(exc)
        rc = lib().pasif_create(C.byref(self._h), p, exc.shape[0], exc.shape[1], int(sampleRate_), int(numsteps_), self.dtype, dev)
        if rc != 0:
            raise PasifError(lib().pasif_last_error(None).decode())
        compat = os.environ.get("PASIF_COMPAT")
        if compat:
            self._setOption("compat", compat)

    def close(self):
        if getattr(self, "_h", None) and self._h.value:
            lib().pasif_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, rc):
        if rc != 0:
            raise PasifError(lib().pasif_last_error(self._h).decode())

    # ---- forward / adjoint setters
    def _set_k(self, fn, n_, values_, indices_, indptr_):
        v, pv = _d(values_)
        i, pi = _u(indices_)
        p, pp = _u(indptr_)
        self._check(fn(self._h, int(n_[0]), int(n_[1]), pv, pi, pp, v.size))

    def _set_t(self, fn, order, n_, values_, indices_):
        d, pd_ = _u(list(n_))
        v, pv = _d(values_)
        i, pi = _u(indices_)
        if i.size != order * v.size:
            raise ValueError("COO tensor needs `order` indices per value")
        self._check(fn(self._h, pd_, pv, pi, v.size))

    def _set_v(self, fn, vec):
        v, pv = _d(vec)
        self._check(fn(self._h, pv, v.size))

    def _setFwdK(self, n_, values_, indices_, indptr_): self._set_k(lib().pasif_set_fwd_k, n_, values_, indices_, indptr_)
    def _setFwdGamma(self, n_, values_, indices_): self._set_t(lib().pasif_set_fwd_gamma, 3, n_, values_, indices_)
    def _setFwdLambda(self, n_, values_, indices_): self._set_t(lib().pasif_set_fwd_lambda, 4, n_, values_, indices_)
    def _setFwdPsi(self, n_, values_, indices_): self._set_t(lib().pasif_set_fwd_psi, 5, n_, values_, indices_)
    def _setFwdForcePattern(self, forcePattern_): self._set_v(lib().pasif_set_fwd_force, forcePattern_)
    def _setFwdInitialConditions(self, initialConditions_): self._set_v(lib().pasif_set_fwd_ic, initialConditions_)
    def _allocateSystemOnDevice(self): self._check(lib().pasif_alloc_system(self._h))
    def _setBwdK(self, n_, values_, indices_, indptr_): self._set_k(lib().pasif_set_bwd_k, n_, values_, indices_, indptr_)
    def _setBwdGamma(self, n_, values_, indices_): self._set_t(lib().pasif_set_bwd_gamma, 3, n_, values_, indices_)
    def _setBwdLambda(self, n_, values_, indices_): self._set_t(lib().pasif_set_bwd_lambda, 4, n_, values_, indices_)
    def _setBwdPsi(self, n_, values_, indices_): self._set_t(lib().pasif_set_bwd_psi, 5, n_, values_, indices_)
    def _setBwdForcePattern(self, forcePattern_): self._set_v(lib().pasif_set_bwd_force, forcePattern_)
    def _setBwdInitialConditions(self, initialConditions_): self._set_v(lib().pasif_set_bwd_ic, initialConditions_)
    def _allocateAdjointOnDevice(self): self._check(lib().pasif_alloc_adjoint(self._h))

    # ---- options
    def _loadExcitationsSet(self, excitationSet_, sampleRate_):
        exc = np.ascontiguousarray(np.asarray(excitationSet_, dtype=np.float64))
        if exc.ndim != 2:
            raise ValueError("Excitations vectors are not of the same size")
        keep, p = _d(exc)
        self._check(lib().pasif_load_excitations(self._h, p, exc.shape[0], exc.shape[1], int(sampleRate_)))
        return 0

    def _setInterpolationMatrix(self, interpolationMatrix_, interpolationWindowSize_):
        v, pv = _d(interpolationMatrix_)
        self._check(lib().pasif_set_interp(self._h, pv, v.size, int(interpolationWindowSize_)))

    def _setInterpolationLinear(self, interpolationWindowSize_):
        """additive: linear 2 x W interpolation table generated on the device"""
        self._check(lib().pasif_set_interp_linear(self._h, int(interpolationWindowSize_)))

    def _setModulationBuffer(self, modulationBuffer_): self._set_v(lib().pasif_set_modulation, modulationBuffer_)

    def _setOption(self, key, value): self._check(lib().pasif_set_option(self._h, str(key).encode(), str(value).encode()))

    def _query(self, key):
        v = C.c_double()
        if lib().pasif_query(self._h, key.encode(), C.byref(v)) != 0:
            raise KeyError(key)
        return v.value

    # ---- solvers (NumPy arrays instead of Python lists; element order identical to the reference)
    def _getAmplitudes(self, fetch=True):
        if not fetch:
            self._check(lib().pasif_get_amplitudes(self._h, None, 0))
            return None
        out = np.empty(int(self._query("n_fwd") * self._query("n_excitations")), dtype=np.float64)
        self._check(lib().pasif_get_amplitudes(self._h, out.ctypes.data_as(_dp), out.size))
        return out

    def _getTrajectory(self, saveSteps_=1):
        out = np.empty(int(lib().pasif_trajectory_size(self._h, int(saveSteps_))), dtype=np.float64)
        self._check(lib().pasif_get_trajectory(self._h, int(saveSteps_), out.ctypes.data_as(_dp), out.size))
        return out

    def _getGradient(self, save=0, fetch=True):
        if not fetch:
            self._check(lib().pasif_get_gradient(self._h, int(save), None, 0))
            return None
        out = np.empty(int(self._query("n_adj") * self._query("n_excitations")), dtype=np.float64)
        self._check(lib().pasif_get_gradient(self._h, int(save), out.ctypes.data_as(_dp), out.size))
        return out

    def _getGradientReduced(self, save=0):
        out = np.empty(int(self._query("n_adj")), dtype=np.float64)
        self._check(lib().pasif_get_gradient_reduced(self._h, int(save), out.ctypes.data_as(_dp), out.size))
        return out

    def _getGradientReducedDevice(self, device_ptr, save=0):
        self._check(lib().pasif_get_gradient_reduced_device(self._h, int(save), C.c_void_p(int(device_ptr)), int(self._query("n_adj"))))

    def _reduceLastGradientDevice(self, device_ptr):
        """sum over excitations of the adjoint state the last `_getGradient` left on the device -> device pointer"""
        self._check(lib().pasif_reduce_last_gradient_device(self._h, C.c_void_p(int(device_ptr)), int(self._query("n_adj"))))

    def emit_patterns(self):
        need = C.c_uint64()
        self._check(lib().pasif_emit_patterns(self._h, None, 0, C.byref(need)))
        buf = C.create_string_buffer(need.value)
        self._check(lib().pasif_emit_patterns(self._h, buf, need.value, C.byref(need)))
        return buf.value.decode()

    def jit_log(self, arch="sm_100a"):
        """NVRTC + ptxas -v logs of the run-time specialised kernels of the loaded classes (no device needed)."""
        need = C.c_uint64()
        self._check(lib().pasif_debug_jit_log(self._h, arch.encode(), None, 0, C.byref(need)))
        buf = C.create_string_buffer(need.value)
        self._check(lib().pasif_debug_jit_log(self._h, arch.encode(), buf, need.value, C.byref(need)))
        return buf.value.decode()


def fma_peak(device=0, dtype="f64"):
    v = C.c_double()
    if lib().pasif_fma_peak(int(device), {"f64": 0, "f32": 1, "dmma": 2, "tf32_mma": 3, "f32x2": 4}[dtype], C.byref(v)) != 0:
        raise PasifError(lib().pasif_last_error(None).decode())
    return v.value


def debug_gemm(kd, x, dtype="f32", mode=0, device=0):
    """y[p][m] = sum_k kd[m][k] x[p][k] through one of the dense-operator path's GEMM kernels (see pasif_debug_gemm)."""
    kd, pk = _d(kd)
    x, px = _d(x)
    n, P = kd.shape[0], x.shape[0]
    assert kd.shape == (n, n) and x.shape == (P, n)
    y = np.empty((P, n), dtype=np.float64)
    if lib().pasif_debug_gemm(int(device), {"f64": 0, "f32": 1}[dtype], int(mode), n, P, pk, px, y.ctypes.data_as(_dp)) != 0:
        raise PasifError(lib().pasif_last_error(None).decode())
    return y


def host_engine_class(dtype="f64"):
    """Driver class whose instances are host-only (device = -1): structural analysis and trait emission without a GPU."""
    class HostEngine(Engine):
        def __init__(self, excitationSet_, sampleRate_, numsteps_, dCompute_=False, dSystem_=False, dSolver_=False, GPUId_=0):
            super().__init__(excitationSet_, sampleRate_, numsteps_, dtype_=dtype, device=-1)
    return HostEngine


def engine_class(dtype="f64", compat=None):
    class _E(Engine):
        def __init__(self, excitationSet_, sampleRate_, numsteps_, dCompute_=False, dSystem_=False, dSolver_=False, GPUId_=0):
            super().__init__(excitationSet_, sampleRate_, numsteps_, GPUId_=GPUId_, dtype_=dtype)
            if compat:
                self._setOption("compat", compat)
    return _E
