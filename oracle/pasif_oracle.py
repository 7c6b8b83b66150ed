"""CPU ORACLE -- TEST INFRASTRUCTURE ONLY.  Not imported by the product (`pasif_b200/`); used by `tests/`,
`__graft_entry__.smoke()` and the `cpu_baseline` / `--impl reference` legs of `bench.py` as the checker.

NumPy restatement of the reference's RK4 forward / adjoint hot path, taking exactly what crosses the
pybind11 boundary `PASIfgpu.__GpuDriver` (reference `src/PASIfgpu.cpp:22-105`).  Vectorised over the
excitation axis; every routine cites the reference lines it restates.

Parity pin: validated against golden vectors produced by running the compiled, unmodified reference on a
B200 (`tests/golden/ref_*.npz`, generator `tests/golden/gen_reference_golden.py`).  The reference's own
repository holds no golden vectors for this path; its only known-answer cases
(`src/utest.cu:220-296`) are checked in `tests/test_oracle.py`.

Two modes for the gradient:
  compat="reference"  literal schedule of `__GpuDriver::_getGradient` incl. quirks A1-A5
  compat="clean"      exact checkpoints, all chunks, no aliasing: the mathematically intended sweep

Call-order state (quirk A7, `call_order = "reference"`, the default): the driver's forward state `d_fwd_Q` is reset to the
initial conditions only AFTER getAmplitudes / getTrajectory (src/__GpuDriver.cu:447, 482) and whenever the system is
(re)allocated (:1183-1186); `_getGradient` starts its forward sweep from whatever it holds (:522-523) and leaves the end state
of the last recomputed chunk behind (:565); the adjoint initial-condition buffer is overwritten with the running adjoint
state (:743), so a second `_getGradient` without `setJacobian` continues from the previous result.  `call_order =
"independent"` makes every call start from the initial conditions.
"""
from __future__ import annotations

import math
import numpy as np


def update_slope(q, m, dt):
    """`updateSlope` kernel (reference src/kernels.cu:201-210): rki = q + dt * rk."""
    return q + dt * m


def integrate(q, m1, m2, m3, m4, h6):
    """`integrate` kernel (reference src/kernels.cu:218-229): q += h6 * (rk1 + 2 rk2 + 2 rk3 + rk4)."""
    return q + h6 * (m1 + 2 * m2 + 2 * m3 + m4)


class _Tensor:
    """COO tensor of order 3/4/5 as handed over by the front end: flat indices [out, in1, in2, ...] per nnz
    (reference `src/helpers.cu:266-652`; index roles from `src/kernels.cu:52,56,61`)."""

    def __init__(self, order, dims, values, indices, dtype):
        self.order = order
        self.dims = [int(d) for d in dims]
        self.val = np.asarray(values, dtype=np.float64).astype(dtype)
        idx = np.asarray(indices, dtype=np.int64).reshape(-1, order) if len(values) else np.zeros((0, order), np.int64)
        self.idx = idx
        self.nnz = len(self.val)


class _System:
    def __init__(self):
        self.n = 0
        self.K = None          # (data, indices, indptr)
        self.G = self.L = self.P = None
        self.F_idx = np.zeros(0, np.int64)
        self.F_val = None
        self.IC = None


class PasifOracle:
    """Mirror of `__GpuDriver` (reference `src/__GpuDriver.cu`) on the CPU."""

    def __init__(self, excitations, sample_rate, numsteps, dtype=np.float64):
        self.dtype = np.dtype(dtype).type
        self.fwd = _System()
        self.bwd = _System()
        self.interp = None
        self.numsteps = int(numsteps)
        self.call_order = "reference"
        self._q_cur = None                     # forward state left behind by the last _getGradient ([E][n]); None = IC
        self._lam_cur = None                   # adjoint state carried into the next _getGradient ([E][n_adj]); None = IC
        self.load_excitations(excitations, sample_rate)

    # ---- __GpuDriver::_loadExcitationsSet / setTimesteps  (src/__GpuDriver.cu:343-381, 632-636)
    def load_excitations(self, excitations, sample_rate):
        exc = np.asarray(excitations, dtype=np.float64)
        if exc.ndim != 2:
            raise ValueError("excitations must be rectangular [E][L]")
        self.exc = np.ascontiguousarray(exc.astype(self.dtype))       # (reel)sample
        self.E, self.L = self.exc.shape
        t = self.dtype
        self.h = t(1.0 / float(sample_rate))
        self.h2 = t(float(self.h) / 2.0)
        self.h6 = t(float(self.h) / 6.0)

    # ---- setters (src/__GpuDriver.cu:197-330)
    def _set(self, sysm, K, G, L, P, F, IC):
        t = self.dtype
        shape, data, indices, indptr = K
        sysm.n = int(shape[0])
        sysm.K = (np.asarray(data, dtype=np.float64).astype(t), np.asarray(indices, np.int64), np.asarray(indptr, np.int64))
        sysm.G = _Tensor(3, G[0], G[1], G[2], t)
        sysm.L = _Tensor(4, L[0], L[1], L[2], t)
        sysm.P = _Tensor(5, P[0], P[1], P[2], t) if P is not None else _Tensor(5, [1] * 5, [], [], t)
        Fd = np.asarray(F, dtype=np.float64).astype(t)
        keep = np.abs(Fd) > np.finfo(t).eps                 # COOVector ctor, src/helpers.cu:674-681 (quirk F3)
        sysm.F_idx = np.nonzero(keep)[0]
        sysm.F_val = Fd[keep]
        sysm.IC = np.asarray(IC, dtype=np.float64).astype(t)

    def set_forward(self, K, G, L, P, F, IC):
        self._set(self.fwd, K, G, L, P, F, IC)
        self._q_cur = None                     # allocateDeviceSystemStatesVector: d_fwd_Q <- IC (src/__GpuDriver.cu:1183-1186)

    def set_adjoint(self, K, G, L, P, F, IC):
        self._set(self.bwd, K, G, L, P, F, IC)
        self._lam_cur = None                   # allocateDeviceAdjointStatesVector: d_bwd_QinitCond <- IC (:1363-1367)

    def _take_q(self):
        """forward state a sweep starts from: the driver's current one (quirk A7) or the initial conditions"""
        if self.call_order == "reference" and self._q_cur is not None and self._q_cur.shape == (self.E, self.fwd.n):
            return np.ascontiguousarray(self._q_cur, dtype=self.dtype)
        return None

    # ---- __GpuDriver::_setInterpolationMatrix (src/__GpuDriver.cu:383-398)
    def set_interpolation(self, table_flat, window):
        self.interp = (np.asarray(table_flat, dtype=np.float64).astype(self.dtype), int(window))

    # ---- derivatives(): cusparseSpMM + SpTdV  (src/__GpuDriver.cu:785-838, src/kernels.cu:21-63)
    def _deriv(self, s, X, Pst):
        """X: [E, n] differentiated state; Pst: [E, n_setpoint] forward ("setpoint") state.  All Lambda terms are
        applied (the reference's SpTdV drops/duplicates cubic terms once nnzG+nnzL+nnzP > 128, quirk F4)."""
        data, indices, indptr = s.K
        Y = np.zeros_like(X)
        for r in range(s.n):                                  # Y = K.X, row by row in CSR order
            for k in range(indptr[r], indptr[r + 1]):
                Y[:, r] += data[k] * X[:, indices[k]]
        G, L, P = s.G, s.L, s.P
        for k in range(G.nnz):
            o, a, b = G.idx[k]
            Y[:, o] += G.val[k] * X[:, a] * Pst[:, b]
        for k in range(L.nnz):
            o, a, b, c = L.idx[k]
            Y[:, o] += L.val[k] * X[:, a] * Pst[:, b] * Pst[:, c]
        for k in range(P.nnz):
            o, a, b, c, d = P.idx[k]
            Y[:, o] += P.val[k] * X[:, a] * Pst[:, b] * Pst[:, c] * Pst[:, d]
        return Y

    # ---- modterpolator(): applyForces / interpolateForces_fwd  (src/__GpuDriver.cu:840-890, src/kernels.cu:71-126)
    def _force(self, Y, step, half):
        s = self.fwd
        if len(s.F_idx) == 0:
            return
        if self.interp is None:
            if 0 < step < self.L:                             # quirk F1; `half` ignored (F2)
                Y[:, s.F_idx] += s.F_val[None, :] * self.exc[:, step][:, None]
        else:
            IM, W = self.interp
            u = 2 * step + (1 if half else 0)
            j = u & (W - 1)
            o = u >> int(math.log2(W))
            if 0 < o < self.L:
                e0 = self.exc[:, o]
                # exc[o+1] at o = L-1 reads the next excitation's first sample in the reference's flat buffer
                # (out of bounds for the last one, quirk F6): restated on the flat [e][t] array, 0 past the end.
                flat = self.exc.reshape(-1)
                pos = np.arange(self.E) * self.L + o + 1
                e1 = np.where(pos < flat.size, flat[np.minimum(pos, flat.size - 1)], self.dtype(0))
                val = IM[W + j] * e0 + IM[j] * e1
                Y[:, s.F_idx] += s.F_val[None, :] * val[:, None]

    # ---- fwdStep()  (src/__GpuDriver.cu:693-713; updateSlope/integrate src/kernels.cu:201-229)
    def _fwd_step(self, Q, step):
        s = self.fwd
        h, h2, h6 = self.h, self.h2, self.h6
        m1 = self._deriv(s, Q, Q)
        self._force(m1, step, False)
        mi = update_slope(Q, m1, h2)
        m2 = self._deriv(s, mi, mi)
        self._force(m2, step, True)
        mi = update_slope(Q, m2, h2)
        m3 = self._deriv(s, mi, mi)
        self._force(m3, step, True)
        mi = update_slope(Q, m3, h)
        m4 = self._deriv(s, mi, mi)
        self._force(m4, step + 1, False)
        return integrate(Q, m1, m2, m3, m4, h6)

    def _ic(self, s):
        return np.tile(s.IC[None, :], (self.E, 1)).astype(self.dtype)     # extendTheVector, src/helpers.cu:808-816

    # ---- _getAmplitudes  (src/__GpuDriver.cu:416-461)
    def get_amplitudes(self):
        q0 = self._take_q()
        Q = q0.copy() if q0 is not None else self._ic(self.fwd)
        for t in range(self.numsteps):
            Q = self._fwd_step(Q, t)
        self._q_cur = None                                        # resetStatesVectors after the D2H copy (:447)
        return Q.reshape(-1).copy()

    # ---- _getTrajectory  (src/__GpuDriver.cu:463-500, save rule :669-687)
    def get_trajectory(self, save_steps=1):
        n = self.fwd.n * self.E
        n_slots = (n * self.numsteps) // save_steps // n if n else 0
        out = np.zeros((n_slots, self.E, self.fwd.n), dtype=self.dtype)
        q0 = self._take_q()
        Q = q0.copy() if q0 is not None else self._ic(self.fwd)
        self._q_cur = None                                        # resetStatesVectors (:482)
        k = 0
        for t in range(self.numsteps):
            Q = self._fwd_step(Q, t)
            if t % save_steps == 0:
                if k < n_slots:                                # the reference would write past its buffer otherwise
                    out[k] = Q
                k += 1
        return out.reshape(-1)

    # ---- bwdStep()  (src/__GpuDriver.cu:746-783); the adjoint ODE is unforced (A6)
    def _bwd_stage_sweep(self, lam, q_a, q_half, q_b):
        s = self.bwd
        h, h2, h6 = self.h, self.h2, self.h6
        m1 = self._deriv(s, lam, q_a)
        mi = update_slope(lam, m1, h2)
        m2 = self._deriv(s, mi, q_half)
        mi = update_slope(lam, m2, h2)
        m3 = self._deriv(s, mi, q_half)
        mi = update_slope(lam, m3, h)
        m4 = self._deriv(s, mi, q_b)
        return integrate(lam, m1, m2, m3, m4, h6)

    def schedule(self):
        """numSetpoints / chunkSize / lastChunkSize exactly as src/__GpuDriver.cu:506-508 (reel arithmetic)."""
        n = self.numsteps
        nsp = int(math.sqrt(n))
        t = self.dtype
        c = int(math.ceil(float(t(t(n) / t(nsp)))))
        last = c - (nsp * c) % n
        return nsp, c, last

    # ---- _getGradient  (src/__GpuDriver.cu:502-587) + backwardRungeKutta (:715-744)
    def get_gradient(self, compat="reference", save=0):
        if compat == "clean":
            return self._gradient_clean()
        n = self.numsteps
        nsp, c, last = self.schedule()
        Nf = self.fwd.n
        slots = nsp + c - 2
        T = np.zeros((slots + 2, self.E, Nf), dtype=self.dtype)   # +2: A8 (one-past-the-end write) made defined
        carry = self.call_order == "reference"
        q0 = self._take_q() if carry else None
        Q = q0.copy() if q0 is not None else self._ic(self.fwd)   # A7: the sweep starts from the current d_fwd_Q (:522-523)
        k = 0
        for t in range(n):                                        # forward sweep, save AFTER step t when t % c == 0 (A1)
            Q = self._fwd_step(Q, t)
            if t % c == 0:
                T[k] = Q
                k += 1
        lam = self._lam_cur.copy() if (carry and self._lam_cur is not None and self._lam_cur.shape == (self.E, self.bwd.n)) else self._ic(self.bwd)   # A7 (:743, :1077)
        S = None                                                  # scratch slot frozen by the once-captured graph (A3)
        for sp in range(nsp - 1, 0, -1):                          # chunk 0 is never processed (A2)
            start = sp * c
            if sp == nsp - 1:
                end, top = n, sp + (last - 1)
            else:
                end, top = start + c, sp + (c - 1)
            Q = T[sp].copy()
            for i in range(end - start):                          # recompute; step index restarts at `start` (A1)
                Q = self._fwd_step(Q, start + i)
                T[sp + i] = Q
            if S is None:
                S = top
            cur = top
            for _ in range(end - start):
                T[S] = T[cur]                                     # request_setpoint  (src/kernels.cu:162-171)
                qa = T[S].copy()
                T[S] = self.dtype(.5) * (T[cur] + T[cur - 1])     # half_setpoint     (:173-184)  (A5; A4 at cur == sp)
                qh = T[S].copy()
                T[S] = T[cur - 1]                                 # previous_setpoint (:186-195)
                qb = T[S].copy()
                lam = self._bwd_stage_sweep(lam, qa, qh, qb)
                cur -= 1
            if sp == save:
                break
        if carry:
            self._q_cur, self._lam_cur = Q.copy(), lam.copy()     # d_fwd_Q after the last recompute; d_bwd_QinitCond <- lambda
        return lam.reshape(-1).copy()

    def _gradient_clean(self):
        """Exact-checkpoint sweep: q_k stored for every k, adjoint RK4 step k uses q_{k+1}, (q_{k+1}+q_k)/2, q_k,
        for k = n-1 ... 0 (all chunks)."""
        n = self.numsteps
        Q = self._ic(self.fwd)
        traj = [Q]
        for t in range(n):
            Q = self._fwd_step(Q, t)
            traj.append(Q)
        lam = self._ic(self.bwd)
        half = self.dtype(.5)
        for k in range(n - 1, -1, -1):
            lam = self._bwd_stage_sweep(lam, traj[k + 1], half * (traj[k + 1] + traj[k]), traj[k])
        return lam.reshape(-1).copy()


# --------------------------------------------------------------------------------------------------------------
# front-end hand-off recorder: a stand-in for `PASIfgpu.__GpuDriver` that captures what PASIf.py sends down
# --------------------------------------------------------------------------------------------------------------

class RecordingDriver:
    """Same constructor + 21 `_methods` as the pybind class (src/PASIfgpu.cpp:26-104); stores the arguments and
    forwards solver calls to a PasifOracle.  `dtype` class attribute selects the emulated `reel`."""
    dtype = np.float64
    compat = "reference"
    call_order = "reference"   # quirk A7 (module docstring); "independent": every call starts from the initial conditions
    oracle_cls = None      # default: PasifOracle (NumPy); pasif_oracle_c.CRecordingDriver swaps in the C oracle

    def __init__(self, excitationSet_, sampleRate_, numsteps_, dCompute_=False, dSystem_=False, dSolver_=False, GPUId_=0):
        self.rec = {}
        cls = self.oracle_cls or PasifOracle
        self._oracle = cls(np.asarray(excitationSet_, dtype=np.float64), sampleRate_, numsteps_, self.dtype)
        self._oracle.call_order = self.call_order
        self._psi_f = None
        self._psi_b = None

    def _setFwdK(self, n_, values_, indices_, indptr_):
        self.rec["fwdK"] = (tuple(int(x) for x in n_), np.array(values_, np.float64), np.array(indices_, np.int64), np.array(indptr_, np.int64))

    def _setFwdGamma(self, n_, values_, indices_):
        self.rec["fwdGamma"] = (list(n_), list(values_), list(indices_))

    def _setFwdLambda(self, n_, values_, indices_):
        self.rec["fwdLambda"] = (list(n_), list(values_), list(indices_))

    def _setFwdPsi(self, n_, values_, indices_):
        self.rec["fwdPsi"] = (list(n_), list(values_), list(indices_))

    def _setFwdForcePattern(self, forcePattern_):
        self.rec["fwdF"] = np.array(forcePattern_, np.float64)

    def _setFwdInitialConditions(self, initialConditions_):
        self.rec["fwdIC"] = np.array(initialConditions_, np.float64)

    def _allocateSystemOnDevice(self):
        r = self.rec
        self._oracle.set_forward(r["fwdK"], r["fwdGamma"], r["fwdLambda"], r.get("fwdPsi"), r["fwdF"], r["fwdIC"])

    def _setBwdK(self, n_, values_, indices_, indptr_):
        self.rec["bwdK"] = (tuple(int(x) for x in n_), np.array(values_, np.float64), np.array(indices_, np.int64), np.array(indptr_, np.int64))

    def _setBwdGamma(self, n_, values_, indices_):
        self.rec["bwdGamma"] = (list(n_), list(values_), list(indices_))

    def _setBwdLambda(self, n_, values_, indices_):
        self.rec["bwdLambda"] = (list(n_), list(values_), list(indices_))

    def _setBwdPsi(self, n_, values_, indices_):
        self.rec["bwdPsi"] = (list(n_), list(values_), list(indices_))

    def _setBwdForcePattern(self, forcePattern_):
        self.rec["bwdF"] = np.array(forcePattern_, np.float64)

    def _setBwdInitialConditions(self, initialConditions_):
        self.rec["bwdIC"] = np.array(initialConditions_, np.float64)

    def _allocateAdjointOnDevice(self):
        r = self.rec
        self._oracle.set_adjoint(r["bwdK"], r["bwdGamma"], r["bwdLambda"], r.get("bwdPsi"), r["bwdF"], r["bwdIC"])

    def _loadExcitationsSet(self, excitationSet_, sampleRate_):
        self._oracle.load_excitations(np.asarray(excitationSet_, dtype=np.float64), sampleRate_)
        return 0

    def _setInterpolationMatrix(self, interpolationMatrix_, interpolationWindowSize_):
        self.rec["interp"] = (np.array(interpolationMatrix_, np.float64), int(interpolationWindowSize_))
        self._oracle.set_interpolation(*self.rec["interp"])

    def _setModulationBuffer(self, modulationBuffer_):
        self.rec["modulation"] = np.array(modulationBuffer_, np.float64)     # never read by any kernel in the reference

    def _setOption(self, key, value):
        if key == "compat":
            self.compat = value
        elif key == "call_order":
            self.call_order = self._oracle.call_order = value
        else:
            raise KeyError(key)

    def _getAmplitudes(self):
        return [float(x) for x in self._oracle.get_amplitudes()]

    def _getTrajectory(self, saveSteps_=1):
        return [float(x) for x in self._oracle.get_trajectory(saveSteps_)]

    def _getGradient(self, save):
        return [float(x) for x in self._oracle.get_gradient(self.compat, save)]
