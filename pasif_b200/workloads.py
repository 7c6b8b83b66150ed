"""Synthetic workloads of the RK4 forward / adjoint hot path.

Every workload is described with plain Python/NumPy data (no front-end classes) so
that the SAME definition can be fed to

  * the unmodified reference front end (``/root/reference/PASIf.py``) when golden
    vectors are generated (``tests/golden/gen_reference_golden.py``),
  * this package's front end (``pasif_b200.frontend.PASIf``),
  * the CPU oracle (``oracle/pasif_oracle.py``) through ``assemble_*`` below.

System definitions follow the reference's examples:
  * ``ex1dof``   -- ``examples/FullInterfaceTesting_1dof.py:44-82`` (forward) and ``:141-205`` (6-state Jacobian)
  * ``duffing``  -- BASELINE config 1: same 4-state layout with a cubic Lambda term on the velocity row
  * ``fit8``     -- ``examples/FullInterfaceTesting.py:48-76`` (8-state forward) and ``:143-230`` (12-state Jacobian)
"""
from __future__ import annotations

import copy
import numpy as np

# --------------------------------------------------------------------------------------
# system descriptions: dict(n, M, B=(val,row,col), K=(val,row,col), Gamma=(dims,val,idx),
#                           Lambda=(dims,val,idx), [Psi=(dims,val,idx)], F, IC)
# --------------------------------------------------------------------------------------


def _sys(n, M, B, K, Gamma, Lambda, F, IC, Psi=None):
    d = dict(n=int(n), M=np.asarray(M, dtype=np.float64),
             B=tuple(np.asarray(a) for a in B), K=tuple(np.asarray(a) for a in K),
             Gamma=(list(Gamma[0]), [float(v) for v in Gamma[1]], [int(i) for i in Gamma[2]]),
             Lambda=(list(Lambda[0]), [float(v) for v in Lambda[1]], [int(i) for i in Lambda[2]]),
             F=np.asarray(F, dtype=np.float64), IC=np.asarray(IC, dtype=np.float64))
    if Psi is not None:
        d["Psi"] = (list(Psi[0]), [float(v) for v in Psi[1]], [int(i) for i in Psi[2]])
    return d


def ex1dof_system(k=1.0, b=1.0):
    """Shipped 1-DOF example: state [x, p2, v, p1]; p1' = x^2 (Gamma probe); zero-valued Lambda entry."""
    S = 4
    return _sys(S, [1., 1., 1., 1.],
                ([-1., -1.], [0, 1], [2, 3]),
                ([k, b], [2, 2], [0, 2]),
                ([S] * 3, [-1.], [3, 0, 0]),
                ([S] * 4, [0.], [0, 0, 0, 0]),
                [0., 0., 1., 0.], np.zeros(S))


def ex1dof_jacobian():
    """Shipped 6-state Jacobian of the 1-DOF example (examples/FullInterfaceTesting_1dof.py:141-205)."""
    A, S = 6, 4
    ic = np.zeros(A)
    ic[0] = 1.
    return _sys(A, np.ones(A),
                ([-1., -1.], [1, 2], [4, 5]),
                ([], [], []),
                ([A, A, S], [], []),
                ([A, A, S, S], [-1., -1.], [4, 0, 0, 0, 5, 0, 1, 1]),
                np.zeros(A), ic,
                Psi=([A, A, S, S, S], [0.], [1, 1, 1, 1, 1]))


def duffing_system(m=1.0, k=1.0e5, b=20.0, lam=1.0e6, F0=1.0e3):
    """BASELINE config 1: m x'' + b x' + k x + lam x^3 = F0 e(t); state [x, p2, v, p1]."""
    S = 4
    return _sys(S, [1., 1., m, 1.],
                ([-1., -1.], [0, 1], [2, 3]),
                ([k, b], [2, 2], [0, 2]),
                ([S] * 3, [-1.], [3, 0, 0]),
                ([S] * 4, [lam], [2, 0, 0, 0]),
                [0., 0., F0, 0.], np.zeros(S))


def quad4_system(k=3.0e4, b=15.0, g=50.0, F0=5.0e2):
    """4-state system whose sparsity pattern differs from `duffing_system` (x*v Gamma term, no Lambda): used to
    exercise several structural classes in one call."""
    S = 4
    return _sys(S, [1., 1., 2., 1.],
                ([-1., -1.], [0, 1], [2, 3]),
                ([k, b], [2, 2], [0, 2]),
                ([S] * 3, [-1., g], [3, 0, 0, 2, 0, 2]),
                ([S] * 4, [], []),
                [0., 0., F0, 0.], np.zeros(S))


def quad4_jacobian():
    """6-state Jacobian variant with an extra Gamma term (pattern differs from `ex1dof_jacobian`)."""
    j = ex1dof_jacobian()
    A, S = 6, 4
    j["Gamma"] = ([A, A, S], [-0.5], [3, 4, 2])
    j["IC"][1] = 0.5
    return j


def fit8_system():
    """8-state 2-DOF system with Gamma + Lambda (examples/FullInterfaceTesting.py:48-76)."""
    S = 8
    return _sys(S, [1., 1., 1., 1., 1., 10., 1., 1.],
                ([-1., -1., -1., -1., 1., 10.], [0, 1, 2, 3, 4, 5], [4, 5, 6, 7, 4, 5]),
                ([6., 10.], [4, 5], [0, 1]),
                ([S] * 3, [-10., -10., -1., -1.], [4, 0, 1, 5, 0, 0, 6, 0, 0, 7, 1, 1]),
                ([S] * 4, [40000.], [5, 1, 1, 1]),
                [0, 0, 0, 0, 0.5, 0, 0, 0], np.zeros(S))


def fit8_jacobian():
    """12-state Jacobian with Gamma/Lambda/Psi (examples/FullInterfaceTesting.py:143-230)."""
    A, S = 12, 8
    ic = np.zeros(A)
    ic[2] = 1.
    ic[3] = 1.
    return _sys(A, np.ones(A),
                ([1., 1., 1., 1.], [4, 5, 6, 7], [0, 1, 2, 3]),
                ([-12.0, -2.0, -1., -1.0], [0, 4, 1, 5], [4, 4, 5, 5]),
                ([A, A, S], [20.0, 2.0, 1., 20., 1., -2.0, -1.0, -0.2, -0.1],
                 [0, 4, 1, 0, 5, 0, 0, 6, 0, 1, 4, 0, 1, 7, 1, 8, 4, 0, 9, 4, 4, 10, 5, 1, 11, 5, 5]),
                ([A, A, S, S], [-120000.0], [1, 5, 1, 1]),
                np.zeros(A), ic,
                Psi=([A, A, S, S, S], [1.], [1, 1, 1, 1, 1]))


def chain_system(ndof, dense=False, seed=3, k=2.0e4, b=8.0, g=40.0, lam=3.0e5, F0=2.0e4):
    """BASELINE config 4: one `ndof`-DOF oscillator chain, state [x_0..x_{n-1}, v_0..v_{n-1}, p2, p1] (N = 2 ndof + 2):
    m_i v_i' = -(K x)_i - (B v)_i - g_i x_i^2 - lam_i x_i^3 + F_i e(t),  x_i' = v_i,  p1' = x_0^2,  p2' = p1
    (the probe pair of examples/FullInterfaceTesting_1dof.py:68-70).  sparse: K tridiagonal (nearest-neighbour springs,
    fixed ends), B diagonal.  dense: K, B random symmetric positive definite with the same spectral scale, so the
    pre-scaled linear operator K' has ~50 % non-zeros and is shared by every excitation replica."""
    rng = np.random.default_rng(seed)
    n = int(ndof)
    N = 2 * n + 2
    p2, p1 = 2 * n, 2 * n + 1
    M = np.ones(N)
    M[n:2 * n] = rng.uniform(0.8, 1.25, size=n)
    bv, br, bc = [-1.0] * n + [-1.0], list(range(n)) + [p2], list(range(n, 2 * n)) + [p1]      # x' = v, p2' = p1
    kv, kr, kc = [], [], []
    if not dense:
        ks = k * rng.uniform(0.8, 1.25, size=n + 1)                                   # springs between neighbours / walls
        for i in range(n):
            kv.append(ks[i] + ks[i + 1]); kr.append(n + i); kc.append(i)
            if i > 0:
                kv.append(-ks[i]); kr.append(n + i); kc.append(i - 1)
            if i < n - 1:
                kv.append(-ks[i + 1]); kr.append(n + i); kc.append(i + 1)
        for i in range(n):
            bv.append(b * rng.uniform(0.8, 1.25)); br.append(n + i); bc.append(n + i)
    else:
        A = rng.standard_normal((n, n)) / np.sqrt(n)
        Kd = k * (A @ A.T + 0.5 * np.eye(n))
        C = rng.standard_normal((n, n)) / np.sqrt(n)
        Bd = b * (C @ C.T + 0.5 * np.eye(n))
        for i in range(n):
            for j in range(n):
                kv.append(Kd[i, j]); kr.append(n + i); kc.append(j)
                bv.append(Bd[i, j]); br.append(n + i); bc.append(n + j)
    gval, gidx = [-1.0], [p1, 0, 0]
    lval, lidx = [], []
    for i in range(n):
        gval.append(g * rng.uniform(0.5, 1.5)); gidx += [n + i, i, i]
        lval.append(lam * rng.uniform(0.5, 1.5)); lidx += [n + i, i, i, i]
    F = np.zeros(N)
    F[n:2 * n] = F0 * rng.uniform(0.2, 1.0, size=n)
    return _sys(N, M, (bv, br, bc), (kv, kr, kc), ([N] * 3, gval, gidx), ([N] * 4, lval, lidx), F, np.zeros(N))


def lattice_system(nx, ny, seed=5, k=2.0e4, b=8.0, g=40.0, lam=3.0e5, F0=2.0e4):
    """nx x ny membrane of unit masses: nearest-neighbour springs in both directions (fixed rim), diagonal damping, on-site
    quadratic and cubic terms, same probe pair as `chain_system`.  State [x (site-major, i * ny + j), v, p2, p1].  A second
    large-component shape for the lane-block split: couplings reach +-1 and +-ny sites."""
    rng = np.random.default_rng(seed)
    n = int(nx) * int(ny)
    N = 2 * n + 2
    p2, p1 = 2 * n, 2 * n + 1
    M = np.ones(N)
    M[n:2 * n] = rng.uniform(0.8, 1.25, size=n)
    bv, br, bc = [-1.0] * n + [-1.0], list(range(n)) + [p2], list(range(n, 2 * n)) + [p1]
    kv, kr, kc = [], [], []
    kx = k * rng.uniform(0.8, 1.25, size=(nx + 1, ny))          # spring between (i-1, j) and (i, j); rows 0 / nx: the rim
    ky = k * rng.uniform(0.8, 1.25, size=(nx, ny + 1))
    for i in range(nx):
        for j in range(ny):
            s_ = i * ny + j
            kv.append(kx[i, j] + kx[i + 1, j] + ky[i, j] + ky[i, j + 1]); kr.append(n + s_); kc.append(s_)
            if i > 0:
                kv.append(-kx[i, j]); kr.append(n + s_); kc.append(s_ - ny)
            if i < nx - 1:
                kv.append(-kx[i + 1, j]); kr.append(n + s_); kc.append(s_ + ny)
            if j > 0:
                kv.append(-ky[i, j]); kr.append(n + s_); kc.append(s_ - 1)
            if j < ny - 1:
                kv.append(-ky[i, j + 1]); kr.append(n + s_); kc.append(s_ + 1)
    for s_ in range(n):
        bv.append(b * rng.uniform(0.8, 1.25)); br.append(n + s_); bc.append(n + s_)
    gval, gidx = [-1.0], [p1, 0, 0]
    lval, lidx = [], []
    for s_ in range(n):
        gval.append(g * rng.uniform(0.5, 1.5)); gidx += [n + s_, s_, s_]
        lval.append(lam * rng.uniform(0.5, 1.5)); lidx += [n + s_, s_, s_, s_]
    F = np.zeros(N)
    F[n:2 * n] = F0 * rng.uniform(0.2, 1.0, size=n)
    return _sys(N, M, (bv, br, bc), (kv, kr, kc), ([N] * 3, gval, gidx), ([N] * 4, lval, lidx), F, np.zeros(N))


def chain_jacobian(system, n_params=4):
    """Adjoint system of `chain_system` for the loss p1(T) = int x_0^2 dt with gradients w.r.t. the first `n_params`
    on-site cubic coefficients (pasif_b200.adjoint.build_adjoint)."""
    from .adjoint import build_adjoint
    N = system["n"]
    loss = np.zeros(N)
    loss[N - 1] = 1.0
    nl = len(system["Lambda"][1])
    return build_adjoint(system, loss, [("Lambda", k) for k in range(min(n_params, nl))])


def jitter(system, rng, rel=0.1):
    """Copy of ``system`` whose B/K/Gamma/Lambda/Psi/F values are scaled by independent U(1-rel, 1+rel)."""
    s = copy.deepcopy(system)

    def jit(v):
        v = np.asarray(v, dtype=np.float64)
        return v * rng.uniform(1.0 - rel, 1.0 + rel, size=v.shape)

    s["B"] = (jit(s["B"][0]), s["B"][1], s["B"][2])
    s["K"] = (jit(s["K"][0]), s["K"][1], s["K"][2])
    for key in ("Gamma", "Lambda", "Psi"):
        if key in s:
            s[key] = (s[key][0], [float(x) for x in jit(s[key][1])], s[key][2])
    s["F"] = jit(s["F"])
    return s


def jitter_nonlinear(system, rng, rel=0.1):
    """Parameter replica of ``system``: the linear operator (M, B, K) is kept bit for bit -- every replica shares K' --
    while Gamma / Lambda / Psi / F values are scaled by independent U(1-rel, 1+rel) (BASELINE config 4: "excitation /
    parameter replicas" of one large-DOF system).  Probe terms (coefficient exactly -1, p1' = x^2) are left alone."""
    s = copy.deepcopy(system)

    def jit(v):
        v = np.asarray(v, dtype=np.float64)
        return np.where(v == -1.0, v, v * rng.uniform(1.0 - rel, 1.0 + rel, size=v.shape))

    for key in ("Gamma", "Lambda", "Psi"):
        if key in s:
            s[key] = (s[key][0], [float(x) for x in jit(s[key][1])], s[key][2])
    s["F"] = jit(s["F"])
    return s


# --------------------------------------------------------------------------------------
# excitations
# --------------------------------------------------------------------------------------

DUFFING_FREQS = (20., 35., 45., 50., 55., 65., 80., 120.)


def sine_excitations(freqs, L, sample_rate, amps=None, phases=None):
    t = np.arange(L, dtype=np.float64) / float(sample_rate)
    freqs = np.asarray(freqs, dtype=np.float64)
    amps = np.ones_like(freqs) if amps is None else np.asarray(amps, dtype=np.float64)
    phases = np.zeros_like(freqs) if phases is None else np.asarray(phases, dtype=np.float64)
    return amps[:, None] * np.sin(2.0 * np.pi * freqs[:, None] * t[None, :] + phases[:, None])


def random_sine_excitations(E, L, sample_rate, seed=0):
    """C2/C3/C5 excitations: e_j = a_j sin(2 pi f_j t + phi_j), f~U(20,2000), a~U(.5,1.5), phi~U(0,2pi)."""
    rng = np.random.default_rng(seed)
    f = rng.uniform(20.0, 2000.0, size=E)
    a = rng.uniform(0.5, 1.5, size=E)
    p = rng.uniform(0.0, 2.0 * np.pi, size=E)
    return sine_excitations(f, L, sample_rate, a, p)


def linear_interpolation_table(W):
    """2 x W table for the reference's `interpolateForces_fwd` (src/kernels.cu:96-126):
    row 0 weights exc[o+1], row 1 weights exc[o]; sub-sample j of W gets (j/W, 1-j/W)."""
    j = np.arange(W, dtype=np.float64)
    return np.stack([j / W, 1.0 - j / W])


# --------------------------------------------------------------------------------------
# adapters
# --------------------------------------------------------------------------------------

def to_frontend_args(systems, mod, with_psi=False):
    """Build the positional argument lists of ``PASIf.setSystems`` / ``setJacobian`` for front-end module ``mod``
    (anything exposing ``coo_tensor``; scipy classes are used for matrices exactly like the reference's examples)."""
    from scipy.sparse import dia_matrix, coo_matrix

    def tensor(t):
        ct = mod.coo_tensor(dimensions_=list(t[0]))
        ct.val = list(t[1])
        ct.indices = list(t[2])
        return ct

    vecM, vecB, vecK, vecG, vecL, vecF, vecIC, vecP = [], [], [], [], [], [], [], []
    for s in systems:
        n = s["n"]
        vecM.append(dia_matrix((np.array(s["M"], dtype=np.float64), [0]), shape=(n, n)))
        vecB.append(coo_matrix((np.array(s["B"][0], dtype=np.float64),
                                (np.array(s["B"][1], dtype=np.int64), np.array(s["B"][2], dtype=np.int64))), shape=(n, n)))
        vecK.append(coo_matrix((np.array(s["K"][0], dtype=np.float64),
                                (np.array(s["K"][1], dtype=np.int64), np.array(s["K"][2], dtype=np.int64))), shape=(n, n)))
        vecG.append(tensor(s["Gamma"]))
        vecL.append(tensor(s["Lambda"]))
        vecF.append(np.array(s["F"], dtype=np.float64))
        vecIC.append(np.array(s["IC"], dtype=np.float64))
        if with_psi:
            vecP.append(tensor(s["Psi"]))
    args = [vecM, vecB, vecK, vecG, vecL, vecF, vecIC]
    if with_psi:
        args.append(vecP)
    return args


# --------------------------------------------------------------------------------------
# golden cases: each one is run through the compiled UNMODIFIED reference on a B200 by
# tests/golden/gen_reference_golden.py; the parity tests rebuild the identical inputs by name.
# All of them keep nnz(Gamma)+nnz(Lambda)+nnz(Psi) <= 128 for the global system, the range in
# which the reference's SpTdV thread map is correct.
# --------------------------------------------------------------------------------------

def golden_case(name):
    """-> dict(systems, jacobians|None, excitations[E][L], sample_rate, numsteps (0 = L), interp=(table, W)|None,
               traj_save_steps|None)"""
    sr = 16000
    if name == "ex1dof":                       # examples/FullInterfaceTesting_1dof.py as shipped
        return dict(systems=[ex1dof_system()], jacobians=[ex1dof_jacobian()],
                    excitations=np.ones((1, 78001)), sample_rate=sr, numsteps=0, interp=None, traj_save_steps=1000)
    if name == "duffing8":                     # BASELINE config 1
        return dict(systems=[duffing_system()], jacobians=None,
                    excitations=sine_excitations(DUFFING_FREQS, 78001, sr), sample_rate=sr, numsteps=0,
                    interp=None, traj_save_steps=1000)
    if name == "fit8_e16":                     # C2 system, 16 excitations
        return dict(systems=[fit8_system()], jacobians=[fit8_jacobian()],
                    excitations=random_sine_excitations(16, 78001, sr, seed=0), sample_rate=sr, numsteps=0,
                    interp=None, traj_save_steps=7800)
    if name == "fit8_c2":                      # BASELINE config 2 (E = 1024)
        return dict(systems=[fit8_system()], jacobians=[fit8_jacobian()],
                    excitations=random_sine_excitations(1024, 78001, sr, seed=0), sample_rate=sr, numsteps=0,
                    interp=None, traj_save_steps=None)
    if name == "fit8x2_jit":                   # two jittered copies -> block-diagonal assembly, per-system values
        # (the reference front end cannot assemble >= 3 Jacobians: PASIf.py:574 concatenates a flattened 1-D
        #  diagonal with the next system's 2-D dia_matrix.data and raises ValueError)
        rng = np.random.default_rng(1)
        return dict(systems=[jitter(fit8_system(), rng) for _ in range(2)],
                    jacobians=[jitter(fit8_jacobian(), rng) for _ in range(2)],
                    excitations=random_sine_excitations(5, 20000, sr, seed=2), sample_rate=sr, numsteps=0,
                    interp=None, traj_save_steps=5000)
    if name == "mixed2":                       # two structurally different systems in one call
        # (same size: the reference front end cannot concatenate systems of different sizes, PASIf.py:508)
        return dict(systems=[duffing_system(k=2.0e4, lam=3.0e5), quad4_system()],
                    jacobians=[ex1dof_jacobian(), quad4_jacobian()],
                    excitations=random_sine_excitations(3, 10000, sr, seed=3), sample_rate=sr, numsteps=0,
                    interp=None, traj_save_steps=2500)
    if name == "ex1dof_interp":                # interpolation table path (W = 8 -> 4 RK4 steps per sample)
        return dict(systems=[ex1dof_system(k=1.0e4, b=10.0)], jacobians=None,
                    excitations=sine_excitations([30., 60., 90., 150.], 2000, sr), sample_rate=4 * sr, numsteps=7900,
                    interp=(linear_interpolation_table(8), 8), traj_save_steps=100)
    if name == "ex1dof_short":                 # numsteps < L, ragged last chunk
        return dict(systems=[ex1dof_system(k=400.0, b=2.0)], jacobians=[ex1dof_jacobian()],
                    excitations=sine_excitations([5., 11.], 6000, sr), sample_rate=sr, numsteps=5000,
                    interp=None, traj_save_steps=500)
    if name == "ex1dof_sq":                    # perfect-square numsteps (reference writes one slot past its buffer, A8)
        return dict(systems=[ex1dof_system(k=400.0, b=2.0)], jacobians=[ex1dof_jacobian()],
                    excitations=sine_excitations([5., 11.], 6000, sr), sample_rate=sr, numsteps=4900,
                    interp=None, traj_save_steps=None)
    if name == "calls_ex1dof_short":           # call-order sequence (run_call_sequence), two excitations
        return dict(golden_case("ex1dof_short"), traj_save_steps=None)
    if name == "calls_ex1dof_e1":              # ... one excitation, so the returned adjoint state can re-seed setJacobian
        return dict(systems=[ex1dof_system(k=400.0, b=2.0)], jacobians=[ex1dof_jacobian()],
                    excitations=sine_excitations([5.], 6000, sr), sample_rate=sr, numsteps=5000,
                    interp=None, traj_save_steps=None)
    raise KeyError(name)


# call-order fixtures: what the reference returns when its solvers are called repeatedly; the last
# three calls re-seed the adjoint from the returned state like examples/FullInterfaceTesting_1dof.py:233-237
CALL_ORDER_CASES = ("calls_ex1dof_short", "calls_ex1dof_e1")


def run_call_sequence(p, case, mod):
    """Drive front end `p` (already holding the case's systems) through the call-order sequence; -> dict of outputs.
    a0 amplitudes | setJacobian | g1, g2 two gradients in a row | a1, a2 two amplitude calls | (one excitation only)
    setJacobian again, then three `_getGradient(save)` calls, each followed by setJacobian with the returned state as the
    adjoint initial condition."""
    import math
    out = {}
    jac = lambda: to_frontend_args(case["jacobians"], mod, with_psi=True)  # noqa: E731
    out["a0"] = np.asarray(p.getAmplitudes(), dtype=np.float64)
    p.setJacobian(*jac())
    out["g1"] = np.asarray(p._getGradient(0), dtype=np.float64)
    out["g2"] = np.asarray(p._getGradient(0), dtype=np.float64)
    out["a1"] = np.asarray(p.getAmplitudes(), dtype=np.float64)
    out["a2"] = np.asarray(p.getAmplitudes(), dtype=np.float64)
    if len(case["excitations"]) == 1 and len(case["systems"]) == 1:
        p.setJacobian(*jac())
        n = case["numsteps"] or len(case["excitations"][0])
        nsp = int(math.sqrt(n))
        for i, save in enumerate((nsp - 1, nsp - 2, nsp - 3)):
            g = np.asarray(p._getGradient(save), dtype=np.float64)
            out[f"s{i}"] = g
            j = copy.deepcopy(case["jacobians"][0])
            j["IC"] = g.copy()
            p.setJacobian(*to_frontend_args([j], mod, with_psi=True))
    return out


GOLDEN_CASES = ("ex1dof", "duffing8", "fit8_e16", "fit8_c2", "fit8x2_jit", "mixed2",
                "ex1dof_interp", "ex1dof_short", "ex1dof_sq")
