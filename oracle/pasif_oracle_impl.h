/* pasif_oracle_impl.h -- body of the plain-C oracle, included twice (REAL = double / float) by pasif_oracle.c.
 * TEST INFRASTRUCTURE ONLY.  Restates, literally and on the GLOBAL system exactly as it crosses the pybind boundary,
 * the reference hot path:
 *   derivative  = cusparseSpMM + SpTdV            reference src/__GpuDriver.cu:785-838, src/kernels.cu:21-63
 *   force       = applyForces / interpolateForces_fwd      src/__GpuDriver.cu:840-890, src/kernels.cu:71-126
 *   step        = fwdStep / bwdStep, updateSlope, integrate src/__GpuDriver.cu:693-713, 746-783, src/kernels.cu:201-229
 *   gradient    = _getGradient + backwardRungeKutta         src/__GpuDriver.cu:502-587, 715-744
 * compiled with -ffp-contract=off: every product and sum is rounded separately (the GPU build fuses some of them;
 * the difference is a few ulp per step and is covered by the parity tolerance).  One trajectory (excitation) at a
 * time, OpenMP over excitations. */

typedef struct {
  int n;
  const REAL* kval; const int* kidx; const int* kptr;
  int nG; const REAL* gval; const int* gidx;
  int nL; const REAL* lval; const int* lidx;
  int nP; const REAL* pval; const int* pidx;
  int nF; const REAL* fval; const int* fidx;
  const REAL* ic;
} NAME(sys_t);

typedef struct {
  const REAL* exc; int E, L;          /* [e][t], already cast to REAL */
  const REAL* table; int W;           /* interpolation table (2 x W) or W = 0 */
  REAL h, h2, h6;
} NAME(ctx_t);

static void NAME(deriv)(const NAME(sys_t)* s, const REAL* X, const REAL* P, REAL* Y) {
  for (int r = 0; r < s->n; ++r) {
    REAL acc = 0;
    for (int k = s->kptr[r]; k < s->kptr[r + 1]; ++k) acc += s->kval[k] * X[s->kidx[k]];
    Y[r] = acc;
  }
  for (int k = 0; k < s->nG; ++k) { const int* i = s->gidx + 3 * k; Y[i[0]] += s->gval[k] * X[i[1]] * P[i[2]]; }
  for (int k = 0; k < s->nL; ++k) { const int* i = s->lidx + 4 * k; Y[i[0]] += s->lval[k] * X[i[1]] * P[i[2]] * P[i[3]]; }
  for (int k = 0; k < s->nP; ++k) { const int* i = s->pidx + 5 * k; Y[i[0]] += s->pval[k] * X[i[1]] * P[i[2]] * P[i[3]] * P[i[4]]; }
}

static void NAME(force)(const NAME(sys_t)* s, const NAME(ctx_t)* c, int e, long step, int half, REAL* Y) {
  if (s->nF == 0) return;
  if (c->W == 0) {
    if (step > 0 && step < c->L) {                                   /* quirk F1; `half` ignored (F2) */
      const REAL x = c->exc[(long)e * c->L + step];
      for (int k = 0; k < s->nF; ++k) Y[s->fidx[k]] += s->fval[k] * x;
    }
  } else {
    const unsigned long u = 2ul * (unsigned long)step + (half ? 1u : 0u);
    const unsigned long j = u & (unsigned long)(c->W - 1);
    unsigned lg = 0; while ((1u << (lg + 1)) <= (unsigned)c->W) ++lg;  /* (uint)log2f(W) */
    const unsigned long o = u >> lg;
    if (o > 0 && o < (unsigned long)c->L) {
      const long pos = (long)e * c->L + (long)o;
      const REAL x0 = c->exc[pos];
      const REAL x1 = (pos + 1 < (long)c->E * c->L) ? c->exc[pos + 1] : (REAL)0;   /* quirk F6 */
      const REAL v = c->table[c->W + j] * x0 + c->table[j] * x1;
      for (int k = 0; k < s->nF; ++k) Y[s->fidx[k]] += s->fval[k] * v;
    }
  }
}

/* one forward RK4 step; work = 5*n REALs */
static void NAME(fwd_step)(const NAME(sys_t)* s, const NAME(ctx_t)* c, int e, long step, REAL* Q, REAL* w) {
  const int n = s->n;
  REAL *m1 = w, *m2 = w + n, *m3 = w + 2 * n, *m4 = w + 3 * n, *mi = w + 4 * n;
  NAME(deriv)(s, Q, Q, m1);   NAME(force)(s, c, e, step, 0, m1);     for (int i = 0; i < n; ++i) mi[i] = Q[i] + c->h2 * m1[i];
  NAME(deriv)(s, mi, mi, m2); NAME(force)(s, c, e, step, 1, m2);     for (int i = 0; i < n; ++i) mi[i] = Q[i] + c->h2 * m2[i];
  NAME(deriv)(s, mi, mi, m3); NAME(force)(s, c, e, step, 1, m3);     for (int i = 0; i < n; ++i) mi[i] = Q[i] + c->h * m3[i];
  NAME(deriv)(s, mi, mi, m4); NAME(force)(s, c, e, step + 1, 0, m4);
  for (int i = 0; i < n; ++i) Q[i] += c->h6 * (m1[i] + 2 * m2[i] + 2 * m3[i] + m4[i]);
}

static void NAME(bwd_step)(const NAME(sys_t)* b, const NAME(ctx_t)* c, REAL* lam, const REAL* qa, const REAL* qh,
                           const REAL* qb, REAL* w) {
  const int n = b->n;
  REAL *m1 = w, *m2 = w + n, *m3 = w + 2 * n, *m4 = w + 3 * n, *mi = w + 4 * n;
  NAME(deriv)(b, lam, qa, m1); for (int i = 0; i < n; ++i) mi[i] = lam[i] + c->h2 * m1[i];
  NAME(deriv)(b, mi, qh, m2);  for (int i = 0; i < n; ++i) mi[i] = lam[i] + c->h2 * m2[i];
  NAME(deriv)(b, mi, qh, m3);  for (int i = 0; i < n; ++i) mi[i] = lam[i] + c->h * m3[i];
  NAME(deriv)(b, mi, qb, m4);
  for (int i = 0; i < n; ++i) lam[i] += c->h6 * (m1[i] + 2 * m2[i] + 2 * m3[i] + m4[i]);
}

/* getAmplitudes / getTrajectory: out_final [E][n]; traj [slot][E][n] (may be NULL).
 * q0 [E][n] or NULL: the forward state the driver currently holds -- the reference starts every sweep from d_fwd_Q as it
 * stands and resets it to the initial conditions only AFTER getAmplitudes / getTrajectory (src/__GpuDriver.cu:428-447,
 * 480-484); NULL = initial conditions. */
int NAME(oracle_forward)(const NAME(sys_t)* s, const NAME(ctx_t)* c, long numsteps, REAL* out_final, long save_steps,
                         REAL* traj, long n_slots, int threads, const REAL* q0) {
  const int n = s->n;
  int err = 0;
#pragma omp parallel for schedule(dynamic, 1) num_threads(threads)
  for (int e = 0; e < c->E; ++e) {
    REAL* Q = (REAL*)malloc(sizeof(REAL) * 6 * (size_t)(n > 0 ? n : 1));
    if (!Q) { err = 1; continue; }
    REAL* w = Q + n;
    memcpy(Q, q0 ? q0 + (size_t)e * n : s->ic, sizeof(REAL) * n);
    long slot = 0;
    for (long t = 0; t < numsteps; ++t) {
      NAME(fwd_step)(s, c, e, t, Q, w);
      if (traj && save_steps > 0 && t % save_steps == 0) {
        if (slot < n_slots) memcpy(traj + ((size_t)slot * c->E + e) * n, Q, sizeof(REAL) * n);
        ++slot;
      }
    }
    if (out_final) memcpy(out_final + (size_t)e * n, Q, sizeof(REAL) * n);
    free(Q);
  }
  return err;
}

/* getGradient.  compat 0 = literal reference schedule (quirks A1-A5), 1 = exact ("clean") sweep.
 * nsp / chunk / last are computed by the caller exactly as src/__GpuDriver.cu:506-508.  out [E][n_adj].
 * Call-order state (quirk A7): q0 [E][nf] / lam0 [E][na] = forward state / adjoint state the driver currently holds (NULL =
 * initial conditions); q_end [E][nf] (may be NULL) receives the forward state the call leaves behind in d_fwd_Q: the end of
 * the last recomputed chunk (src/__GpuDriver.cu:565), or of the sweep when no chunk is processed. */
int NAME(oracle_gradient)(const NAME(sys_t)* f, const NAME(sys_t)* b, const NAME(ctx_t)* c, long numsteps, int compat,
                          long nsp, long chunk, long last, long save, REAL* out, int threads, const REAL* q0,
                          const REAL* lam0, REAL* q_end) {
  const int nf = f->n, na = b->n;
  int err = 0;
#pragma omp parallel for schedule(dynamic, 1) num_threads(threads)
  for (int e = 0; e < c->E; ++e) {
    const long slots = (compat == 0) ? (nsp + chunk + 1) : (numsteps + 1);
    REAL* T = (REAL*)calloc((size_t)slots * (nf > 0 ? nf : 1), sizeof(REAL));
    REAL* Q = (REAL*)malloc(sizeof(REAL) * (6 * (size_t)nf + 6 * (size_t)na + 3 * (size_t)nf + 16));
    if (!T || !Q) { err = 1; free(T); free(Q); continue; }
    REAL* wf = Q + nf;
    REAL* lam = wf + 5 * nf;
    REAL* wb = lam + na;
    REAL* qa = wb + 5 * na; REAL* qh = qa + nf; REAL* qb = qh + nf;
    memcpy(Q, q0 ? q0 + (size_t)e * nf : f->ic, sizeof(REAL) * nf);
    memcpy(lam, lam0 ? lam0 + (size_t)e * na : b->ic, sizeof(REAL) * na);
    if (compat != 0) {
      /* exact sweep: store every state, adjoint step k uses q_{k+1}, (q_{k+1}+q_k)/2, q_k for k = n-1 .. 0 */
      memcpy(T, Q, sizeof(REAL) * nf);
      for (long t = 0; t < numsteps; ++t) { NAME(fwd_step)(f, c, e, t, Q, wf); memcpy(T + (size_t)(t + 1) * nf, Q, sizeof(REAL) * nf); }
      for (long k = numsteps - 1; k >= 0; --k) {
        const REAL* a = T + (size_t)(k + 1) * nf; const REAL* p = T + (size_t)k * nf;
        for (int i = 0; i < nf; ++i) qh[i] = (REAL).5 * (a[i] + p[i]);
        NAME(bwd_step)(b, c, lam, a, qh, p, wb);
      }
    } else {
      long k = 0;
      for (long t = 0; t < numsteps; ++t) {                     /* checkpoint AFTER step t when t % c == 0 (A1) */
        NAME(fwd_step)(f, c, e, t, Q, wf);
        if (t % chunk == 0) { memcpy(T + (size_t)k * nf, Q, sizeof(REAL) * nf); ++k; }
      }
      long S = -1;                                              /* scratch slot frozen by the first graph capture (A3) */
      for (long sp = nsp - 1; sp >= 1; --sp) {                  /* chunk 0 never processed (A2) */
        const long start = sp * chunk;
        long end, top;
        if (sp == nsp - 1) { end = numsteps; top = sp + (last - 1); } else { end = start + chunk; top = sp + (chunk - 1); }
        memcpy(Q, T + (size_t)sp * nf, sizeof(REAL) * nf);
        for (long i = 0; i < end - start; ++i) {                /* recompute with step index restarting at `start` (A1) */
          NAME(fwd_step)(f, c, e, start + i, Q, wf);
          memcpy(T + (size_t)(sp + i) * nf, Q, sizeof(REAL) * nf);
        }
        if (S < 0) S = top;
        long cur = top;
        REAL* TS = T + (size_t)S * nf;
        for (long it = 0; it < end - start; ++it) {
          const REAL* tc = T + (size_t)cur * nf; const REAL* tp = T + (size_t)(cur - 1) * nf;
          for (int i = 0; i < nf; ++i) TS[i] = tc[i];                         /* request_setpoint  */
          memcpy(qa, TS, sizeof(REAL) * nf);
          for (int i = 0; i < nf; ++i) TS[i] = (REAL).5 * (tc[i] + tp[i]);    /* half_setpoint (A5; A4 at cur == sp) */
          memcpy(qh, TS, sizeof(REAL) * nf);
          for (int i = 0; i < nf; ++i) TS[i] = tp[i];                         /* previous_setpoint */
          memcpy(qb, TS, sizeof(REAL) * nf);
          NAME(bwd_step)(b, c, lam, qa, qh, qb, wb);
          --cur;
        }
        if (sp == save) break;
      }
    }
    memcpy(out + (size_t)e * na, lam, sizeof(REAL) * na);
    if (q_end) memcpy(q_end + (size_t)e * nf, Q, sizeof(REAL) * nf);
    free(T); free(Q);
  }
  return err;
}
