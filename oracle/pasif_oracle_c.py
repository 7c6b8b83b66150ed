"""ctypes wrapper of the plain-C oracle (oracle/pasif_oracle.c).  TEST INFRASTRUCTURE ONLY.
Same interface as `pasif_oracle.PasifOracle`; used where the NumPy restatement would be too slow (long time ranges,
many excitations) and as the reported CPU baseline.  Build with `make -C oracle oracle`."""
from __future__ import annotations

import ctypes as C
import math
import os

import numpy as np

from pasif_oracle import PasifOracle, RecordingDriver

HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


def _lib():
    global _LIB
    if _LIB is None:
        path = os.path.join(HERE, "libpasif_oracle.so")
        if not os.path.exists(path):
            import subprocess
            subprocess.check_call(["make", "-C", HERE, "oracle"], stdout=subprocess.DEVNULL)
        _LIB = C.CDLL(path)
    return _LIB


def _sys_struct(real):
    rp, ip = C.POINTER(real), C.POINTER(C.c_int)

    class Sys(C.Structure):
        _fields_ = [("n", C.c_int), ("kval", rp), ("kidx", ip), ("kptr", ip),
                    ("nG", C.c_int), ("gval", rp), ("gidx", ip),
                    ("nL", C.c_int), ("lval", rp), ("lidx", ip),
                    ("nP", C.c_int), ("pval", rp), ("pidx", ip),
                    ("nF", C.c_int), ("fval", rp), ("fidx", ip), ("ic", rp)]

    class Ctx(C.Structure):
        _fields_ = [("exc", rp), ("E", C.c_int), ("L", C.c_int), ("table", rp), ("W", C.c_int),
                    ("h", real), ("h2", real), ("h6", real)]
    return Sys, Ctx


class COracle(PasifOracle):
    """Setters inherited from the NumPy oracle (same rounding of the hand-off to `reel`); solvers run in C."""

    def __init__(self, excitations, sample_rate, numsteps, dtype=np.float64, threads=None):
        super().__init__(excitations, sample_rate, numsteps, dtype)
        self.threads = int(threads or os.cpu_count() or 1)
        self._real = C.c_double if self.dtype == np.float64 else C.c_float
        self._sfx = "f64" if self.dtype == np.float64 else "f32"
        self._Sys, self._Ctx = _sys_struct(self._real)

    def _pack_sys(self, s, keep):
        rp, ip = C.POINTER(self._real), C.POINTER(C.c_int)

        def r(a):
            a = np.ascontiguousarray(a, dtype=self.dtype)
            keep.append(a)
            return a.ctypes.data_as(rp)

        def i(a):
            a = np.ascontiguousarray(a, dtype=np.int32)
            keep.append(a)
            return a.ctypes.data_as(ip)
        data, indices, indptr = s.K
        return self._Sys(s.n, r(data), i(indices), i(indptr),
                         s.G.nnz, r(s.G.val), i(s.G.idx.reshape(-1)),
                         s.L.nnz, r(s.L.val), i(s.L.idx.reshape(-1)),
                         s.P.nnz, r(s.P.val), i(s.P.idx.reshape(-1)),
                         len(s.F_idx), r(s.F_val if s.F_val is not None else []), i(s.F_idx), r(s.IC))

    def _pack_ctx(self, keep):
        rp = C.POINTER(self._real)
        exc = np.ascontiguousarray(self.exc, dtype=self.dtype)
        keep.append(exc)
        if self.interp is not None:
            tab = np.ascontiguousarray(self.interp[0], dtype=self.dtype)
            keep.append(tab)
            tp, W = tab.ctypes.data_as(rp), self.interp[1]
        else:
            tp, W = None, 0
        return self._Ctx(exc.ctypes.data_as(rp), self.E, self.L, tp, W, self._real(self.h), self._real(self.h2), self._real(self.h6))

    def _forward(self, save_steps, n_slots):
        keep = []
        q0 = self._take_q()                      # call-order state (quirk A7), None = initial conditions
        self._q_cur = None                       # getAmplitudes / getTrajectory reset the state afterwards
        s, c = self._pack_sys(self.fwd, keep), self._pack_ctx(keep)
        rp = C.POINTER(self._real)
        out = np.zeros((self.E, self.fwd.n), dtype=self.dtype)
        traj = np.zeros((n_slots, self.E, self.fwd.n), dtype=self.dtype) if n_slots else None
        fn = getattr(_lib(), "oracle_forward_" + self._sfx)
        fn.restype = C.c_int
        rc = fn(C.byref(s), C.byref(c), C.c_long(self.numsteps), out.ctypes.data_as(rp), C.c_long(save_steps),
                traj.ctypes.data_as(rp) if traj is not None else None, C.c_long(n_slots), C.c_int(self.threads),
                q0.ctypes.data_as(rp) if q0 is not None else None)
        if rc:
            raise MemoryError("oracle_forward")
        return out, traj

    def get_amplitudes(self):
        return self._forward(0, 0)[0].reshape(-1)

    def get_trajectory(self, save_steps=1):
        n = self.fwd.n * self.E
        n_slots = (n * self.numsteps) // save_steps // n if n else 0
        traj = self._forward(save_steps, n_slots)[1]
        return traj.reshape(-1) if traj is not None else np.zeros(0, self.dtype)

    def get_gradient(self, compat="reference", save=0):
        keep = []
        f, b, c = self._pack_sys(self.fwd, keep), self._pack_sys(self.bwd, keep), self._pack_ctx(keep)
        rp = C.POINTER(self._real)
        nsp, chunk, last = self.schedule() if self.numsteps > 0 else (0, 1, 1)
        out = np.zeros((self.E, self.bwd.n), dtype=self.dtype)
        carry = compat == "reference" and self.call_order == "reference"
        q0 = self._take_q() if carry else None
        lam0 = self._lam_cur if (carry and self._lam_cur is not None and self._lam_cur.shape == (self.E, self.bwd.n)) else None
        q_end = np.zeros((self.E, self.fwd.n), dtype=self.dtype) if carry else None
        fn = getattr(_lib(), "oracle_gradient_" + self._sfx)
        fn.restype = C.c_int
        rc = fn(C.byref(f), C.byref(b), C.byref(c), C.c_long(self.numsteps), C.c_int(0 if compat == "reference" else 1),
                C.c_long(nsp), C.c_long(chunk), C.c_long(last), C.c_long(save), out.ctypes.data_as(rp), C.c_int(self.threads),
                q0.ctypes.data_as(rp) if q0 is not None else None, lam0.ctypes.data_as(rp) if lam0 is not None else None,
                q_end.ctypes.data_as(rp) if q_end is not None else None)
        if rc:
            raise MemoryError("oracle_gradient")
        if carry:
            self._q_cur, self._lam_cur = q_end, out.copy()
        return out.reshape(-1)


class CRecordingDriver(RecordingDriver):
    """`PASIfgpu.__GpuDriver` stand-in backed by the C oracle."""
    oracle_cls = COracle
