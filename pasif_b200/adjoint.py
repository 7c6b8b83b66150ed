"""Adjoint-system builder: derive the "Jacobian" system that `PASIf.setJacobian`
expects from a forward system, a loss and a parameter list -- what the reference leaves to the caller (hand-typed in
`examples/FullInterfaceTesting.py:147-230`, built by the external `Springtronics` in `examples/test_2DOFgammalambda.py`).

Forward ODE after the front end's pre-scaling (reference PASIf.py:607-618):
    q' = f(q) = K' q + Gamma':(q,q) + Lambda':(q,q,q) + F' e(t),   K' = -M^-1 (B+K), Gamma' = -M^-1 Gamma, ...
Adjoint ODE integrated by the reference in reversed time with positive step (src/__GpuDriver.cu:746-783):
    lam' = J(q)^T lam,  J_ij = df_i/dq_j          -> K_adj[j,i] = K'[i,j]
                                                     Gamma_adj[j,i,b] += Gamma'[i,a,b] (a = j),  += Gamma'[i,b',j] ...
    g_p' = lam_i * d f_i / d theta_p               -> one trailing adjoint variable per parameter
        theta = K'[i,j]            : g' = lam_i q_j            (Gamma_adj entry  [N+p, i, j])
        theta = Gamma'[i,a,b]      : g' = lam_i q_a q_b        (Lambda_adj entry [N+p, i, a, b])
        theta = Lambda'[i,a,b,c]   : g' = lam_i q_a q_b q_c    (Psi_adj entry    [N+p, i, a, b, c])
The result is returned in the USER form of `setJacobian` with M = 1, B = 0: the front end multiplies K, Gamma, Lambda,
Psi by -M^-1 = -1 (PASIf.py:620-633), so every value is stored negated.  Initial condition = dLoss/dq(T).
Gradients are with respect to the PRE-SCALED coefficients (K', Gamma', Lambda').
"""
from __future__ import annotations

import numpy as np


def prescaled(system):
    """(K' as dict (i,j)->v, Gamma' list of (i,a,b,v), Lambda' list of (i,a,b,c,v)) of a `workloads` system dict."""
    n = system["n"]
    minv = -1.0 / np.asarray(system["M"], dtype=np.float64)
    K = {}
    for vals, rows, cols in (system["B"], system["K"]):
        for v, i, j in zip(np.asarray(vals, dtype=np.float64), np.asarray(rows, dtype=np.int64), np.asarray(cols, dtype=np.int64)):
            K[(int(i), int(j))] = K.get((int(i), int(j)), 0.0) + float(v) * minv[int(i)]
    G = []
    gv, gi = system["Gamma"][1], np.asarray(system["Gamma"][2], dtype=np.int64).reshape(-1, 3)
    for v, (i, a, b) in zip(gv, gi):
        G.append((int(i), int(a), int(b), float(v) * minv[int(i)]))
    L = []
    lv, li = system["Lambda"][1], np.asarray(system["Lambda"][2], dtype=np.int64).reshape(-1, 4)
    for v, (i, a, b, c) in zip(lv, li):
        L.append((int(i), int(a), int(b), int(c), float(v) * minv[int(i)]))
    assert all(0 <= i < n for (i, _) in K)
    return K, G, L


def build_adjoint(system, loss_grad, params=()):
    """Jacobian system (dict in the `workloads` format, with "Psi") for `setJacobian`.

    loss_grad : array [n]   dLoss/dq at the final time (e.g. the unit vector of a probe variable)
    params    : sequence of ("K", i, j) | ("Gamma", k) | ("Lambda", k): coefficient of the pre-scaled forward operator
                (k = position in the system's Gamma / Lambda list) whose gradient is wanted; gradient p is adjoint
                variable n + p of the result."""
    n = int(system["n"])
    K, G, L = prescaled(system)
    A = n + len(params)
    kk, gg, ll, pp = {}, {}, {}, {}

    def acc(d, key, v):
        d[key] = d.get(key, 0.0) + v

    for (i, j), v in K.items():
        acc(kk, (j, i), v)
    for (i, a, b, v) in G:
        acc(gg, (a, i, b), v)
        acc(gg, (b, i, a), v)
    for (i, a, b, c, v) in L:
        acc(ll, (a, i, b, c), v)
        acc(ll, (b, i, a, c), v)
        acc(ll, (c, i, a, b), v)
    for p, spec in enumerate(params):
        row = n + p
        if spec[0] == "K":
            acc(gg, (row, int(spec[1]), int(spec[2])), 1.0)
        elif spec[0] == "Gamma":
            i, a, b, _ = G[int(spec[1])]
            acc(ll, (row, i, a, b), 1.0)
        elif spec[0] == "Lambda":
            i, a, b, c, _ = L[int(spec[1])]
            acc(pp, (row, i, a, b, c), 1.0)
        else:
            raise ValueError("unknown parameter kind %r" % (spec[0],))

    def flat(d):
        keys = sorted(d)
        return [-d[k] for k in keys], [int(x) for k in keys for x in k]

    kkeys = sorted(kk)
    ic = np.zeros(A)
    ic[:n] = np.asarray(loss_grad, dtype=np.float64)
    gval, gidx = flat(gg)
    lval, lidx = flat(ll)
    pval, pidx = flat(pp)
    return dict(n=A, M=np.ones(A),
                B=(np.zeros(0), np.zeros(0, np.int64), np.zeros(0, np.int64)),
                K=(np.array([-kk[k] for k in kkeys]), np.array([k[0] for k in kkeys], np.int64),
                   np.array([k[1] for k in kkeys], np.int64)),
                Gamma=([A, A, n], gval, gidx), Lambda=([A, A, n, n], lval, lidx), Psi=([A, A, n, n, n], pval, pidx),
                F=np.zeros(A), IC=ic)
