// large_kernels.cuh -- shared-memory tile kernels for LARGE components (more coupled state variables than the
// register-resident kernels of rk4_kernels.cuh can hold; BASELINE config C4: one 256-DOF chain x 16k excitations).
//
// One CTA integrates one TILE = BT trajectories (consecutive excitations of one system instance) for the whole time
// range of the launch.  Thread r owns state variable (row) r of all BT trajectories:
//   * q[BT], the running stage sum s[BT] and the derivative y[BT] of its row live in registers;
//   * the RK4 stage vector every row needs is published in shared memory as X[row][BTP] (BTP = BT + 1: lanes walk
//     consecutive rows at an odd element stride -> conflict-free), double-buffered so a stage costs ONE barrier;
//   * the row's sparse "program" (ELL slots of K', Gamma', Lambda', Psi' in hand-off order, padded per warp) is read
//     through the read-only path once per stage and amortised over the BT trajectories of the thread -- no atomics
//     (reference SpTdV, src/kernels.cu:21-63), no SpMM launch (src/__GpuDriver.cu:795-807), no per-step launches;
//   * when the linear operator is dense and shared by the whole batch the K'.X product of a stage is a real GEMM
//     [N x N].[N x BT] on the tensor cores (DMMA m8n8k4 in fp64, 3xTF32 m16n8k8 in fp32), see gemm_stage().
// Forward state that the adjoint sweep needs goes to HBM tile-blocked ([step][tile][row][BT]): every thread writes /
// reads whole 32-128 B rows, a tile of a step is one contiguous block.
//
// Arithmetic contract = rk4_kernels.cuh: per row K entries in CSR order, then Gamma, Lambda,
// Psi terms in hand-off order, then F*e; ((val*x_a)*p_b)*p_c.. association; s = m1 + 2 m2 + 2 m3 (+ m4), q += h6*s.
#pragma once

#include <cuda_runtime.h>

namespace pasif_large {

struct Tables {
  int N;                 // state variables of the differentiated system
  int NR;                // N padded to a multiple of 32 (= rows = threads that own a row)
  int NPR;               // padded rows of the set-point vector P (forward: == NR; adjoint: forward system's NR)
  int WK, WG, WL, WP;    // ELL widths (max over rows)
  const int* k_col;      // [WK][NR]
  const int2* g_idx;     // [WG][NR]  (a, b)
  const int4* l_idx;     // [WL][NR]  (a, b, c, -)
  const int4* p_idx;     // [WP][NR]  (a, b, c, d)
  const int4* widths;    // [NR/32]   per-warp slot counts (wk, wg, wl, wp): loops are warp-uniform
  int coef_stride;       // values per instance = (WK + WG + WL + WP + 1) * NR   (last slot: force pattern F')
  int flags;             // dense-operator update kernel: operand indices that coincide in EVERY entry of a table, so one gather
                         // serves two operands: 1 Gamma a == b, 2 Lambda a == b, 4 Lambda b == c (on-site x^2 / x^3 terms)
};

template <typename T>
struct FwdArgs {
  Tables tb;
  const T* vals;         // [n_inst][coef_stride]
  const T* ic;           // [n_inst][NR]
  const T* q_in;         // tile-blocked start state [tile][NR][BT], or nullptr -> ic
  const T* exc;          // gated samples [t][E] (full mode) / [2t+half][E] (half mode), zero rows past the end
  int E, n_inst, tiles_per_inst, n_tiles;
  int t0, t1, half_mode;
  T h, h2, h6;
  T* ckpt;               // [slot][tile][NR][BT]; saved after step t when (t + save_phase) % save_every == 0
  int save_every, save_phase, save_slots;
  T* dense;              // [t - t0][tile][NR][BT]: state after EVERY step (adjoint recompute)
  T* user_traj;          // [slot][E][n_glob] through varmap (getTrajectory)
  int traj_every, traj_slots;
  T* user_final;         // [E][n_glob] (getAmplitudes)
  const int* varmap;     // [n_inst][NR], -1 for padding rows
  int n_glob;
  // dense shared linear operator (GEMM mode) or nullptr: row-major [NR][NR]; lo part only for the fp32 3xTF32 split
  const T* kd_hi;
  const T* kd_lo;
};

template <typename T>
struct BwdArgs {
  Tables tb;             // adjoint pattern; NPR = forward rows
  const T* vals;         // [n_inst][coef_stride]
  const T* ic;           // [n_inst][NR] adjoint initial condition (used when from_ic)
  T* lam;                // tile-blocked adjoint state [tile][NR][BT], in/out
  const T* dense;        // [i][tile][NPR][BT] forward state after step t0 + i
  const T* prev;         // [tile][NPR][BT] forward state standing in for index -1
  int len, alias_local, from_ic;
  int E, n_inst, tiles_per_inst, n_tiles;
  T h, h2, h6;
  T* user_out;           // [E][n_glob] through varmap when non-null (written after the last step of this launch)
  const int* varmap;
  int n_glob;
  const T* kd_hi;
  const T* kd_lo;
};

#define PL_DEV __device__ __forceinline__

template <typename T, int BT>
PL_DEV void load_row(const T* __restrict__ src, T (&v)[BT]) {
  if constexpr (sizeof(T) == 8) {
    const double2* s2 = reinterpret_cast<const double2*>(src);
#pragma unroll
    for (int k = 0; k < BT / 2; ++k) { const double2 t = __ldg(s2 + k); v[2 * k] = t.x; v[2 * k + 1] = t.y; }
  } else {
    const float4* s4 = reinterpret_cast<const float4*>(src);
#pragma unroll
    for (int k = 0; k < BT / 4; ++k) { const float4 t = __ldg(s4 + k); v[4 * k] = t.x; v[4 * k + 1] = t.y; v[4 * k + 2] = t.z; v[4 * k + 3] = t.w; }
  }
}
template <typename T, int BT>
PL_DEV void store_row(T* __restrict__ dst, const T (&v)[BT]) {
  if constexpr (sizeof(T) == 8) {
    double2* d2 = reinterpret_cast<double2*>(dst);
#pragma unroll
    for (int k = 0; k < BT / 2; ++k) d2[k] = make_double2(v[2 * k], v[2 * k + 1]);
  } else {
    float4* d4 = reinterpret_cast<float4*>(dst);
#pragma unroll
    for (int k = 0; k < BT / 4; ++k) d4[k] = make_float4(v[4 * k], v[4 * k + 1], v[4 * k + 2], v[4 * k + 3]);
  }
}

// y = sparse part of the derivative of row r:  sum_K + sum_Gamma + sum_Lambda + sum_Psi   (accumulated INTO y)
template <typename T, int BT, int BTP>
PL_DEV void row_terms(const Tables& tb, const T* __restrict__ v, int r, int4 w, const T* __restrict__ X,
                      const T* __restrict__ P, T (&y)[BT]) {
  const int NR = tb.NR;
  for (int j = 0; j < w.x; ++j) {
    const int col = __ldg(tb.k_col + j * NR + r);
    const T kv = __ldg(v + (size_t)j * NR);
    const T* x = X + col * BTP;
#pragma unroll
    for (int b = 0; b < BT; ++b) y[b] = fma(kv, x[b], y[b]);
  }
  v += (size_t)tb.WK * NR;
  for (int j = 0; j < w.y; ++j) {
    const int2 ix = __ldg(tb.g_idx + j * NR + r);
    const T gv = __ldg(v + (size_t)j * NR);
    const T* xa = X + ix.x * BTP;
    const T* pb = P + ix.y * BTP;
#pragma unroll
    for (int b = 0; b < BT; ++b) y[b] = fma(gv * xa[b], pb[b], y[b]);
  }
  v += (size_t)tb.WG * NR;
  for (int j = 0; j < w.z; ++j) {
    const int4 ix = __ldg(tb.l_idx + j * NR + r);
    const T lv = __ldg(v + (size_t)j * NR);
    const T* xa = X + ix.x * BTP;
    const T* pb = P + ix.y * BTP;
    const T* pc = P + ix.z * BTP;
#pragma unroll
    for (int b = 0; b < BT; ++b) y[b] = fma(lv * xa[b] * pb[b], pc[b], y[b]);
  }
  v += (size_t)tb.WL * NR;
  for (int j = 0; j < w.w; ++j) {
    const int4 ix = __ldg(tb.p_idx + j * NR + r);
    const T pv = __ldg(v + (size_t)j * NR);
    const T* xa = X + ix.x * BTP;
    const T* pb = P + ix.y * BTP;
    const T* pc = P + ix.z * BTP;
    const T* pd = P + ix.w * BTP;
#pragma unroll
    for (int b = 0; b < BT; ++b) y[b] = fma(pv * xa[b] * pb[b] * pc[b], pd[b], y[b]);
  }
}

// ------------------------------------------------------------------------------------------------ tensor-core linear term
// Y[row][b] = sum_k Kd[row][k] * X[k][b] for the whole tile, rows split over the warps of the CTA (32 rows per warp).
// fp64: DMMA  mma.sync.aligned.m8n8k4.row.col.f64   (A = Kd fragment straight from global/L2, B = X from shared)
// fp32: 3xTF32 mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32: Kd = hi + lo split once on the host side of the
//       launch (both parts exactly representable in tf32), X split on the fly; hi*hi + lo*hi + hi*lo, fp32 accumulate.
// The result is written to Yb[row][BTP] in shared memory (the row owners pick it up after a barrier).
template <int BT, int BTP>
PL_DEV void gemm_stage(const double* __restrict__ Kd, const double* /*lo*/, int NR, const double* __restrict__ X,
                       double* __restrict__ Yb) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int g = lane >> 2, t = lane & 3;
  constexpr int NT = BT / 8;                       // n-tiles of 8 trajectories
  double acc[4][NT][2];
#pragma unroll
  for (int m = 0; m < 4; ++m)
#pragma unroll
    for (int n = 0; n < NT; ++n) acc[m][n][0] = acc[m][n][1] = 0.0;
  const double* arow = Kd + (size_t)(warp * 32 + g) * NR + t;      // A fragment: row g of the m-tile, column k0 + t
  for (int k0 = 0; k0 < NR; k0 += 4) {
    double a[4], bfr[NT];
#pragma unroll
    for (int m = 0; m < 4; ++m) a[m] = __ldg(arow + (size_t)m * 8 * NR + k0);
#pragma unroll
    for (int n = 0; n < NT; ++n) bfr[n] = X[(k0 + t) * BTP + n * 8 + g];   // B fragment: k = k0 + t, column g
#pragma unroll
    for (int m = 0; m < 4; ++m)
#pragma unroll
      for (int n = 0; n < NT; ++n)
        asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                     : "+d"(acc[m][n][0]), "+d"(acc[m][n][1])
                     : "d"(a[m]), "d"(bfr[n]));
  }
  // C fragment: row g, columns 2t, 2t+1
#pragma unroll
  for (int m = 0; m < 4; ++m)
#pragma unroll
    for (int n = 0; n < NT; ++n) {
      double* y = Yb + (warp * 32 + m * 8 + g) * BTP + n * 8 + 2 * t;
      y[0] = acc[m][n][0];
      y[1] = acc[m][n][1];
    }
}

PL_DEV unsigned f2tf32(float x) {
  unsigned r;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
  return r;
}

template <int BT, int BTP>
PL_DEV void gemm_stage(const float* __restrict__ Khi, const float* __restrict__ Klo, int NR, const float* __restrict__ X,
                       float* __restrict__ Yb) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int g = lane >> 2, t = lane & 3;
  constexpr int NT = BT / 8;
  float acc[2][NT][4];
#pragma unroll
  for (int m = 0; m < 2; ++m)
#pragma unroll
    for (int n = 0; n < NT; ++n)
#pragma unroll
      for (int c = 0; c < 4; ++c) acc[m][n][c] = 0.f;
  const size_t rbase = (size_t)(warp * 32 + g) * NR + t;
  for (int k0 = 0; k0 < NR; k0 += 8) {
    unsigned ah[2][4], al[2][4], bh[NT][2], bl[NT][2];
#pragma unroll
    for (int m = 0; m < 2; ++m) {
      // a0:(g, t) a1:(g+8, t) a2:(g, t+4) a3:(g+8, t+4) of the 16x8 tile
      const size_t o = rbase + (size_t)m * 16 * NR + k0;
      ah[m][0] = __float_as_uint(__ldg(Khi + o));              al[m][0] = __float_as_uint(__ldg(Klo + o));
      ah[m][1] = __float_as_uint(__ldg(Khi + o + 8 * (size_t)NR));     al[m][1] = __float_as_uint(__ldg(Klo + o + 8 * (size_t)NR));
      ah[m][2] = __float_as_uint(__ldg(Khi + o + 4));          al[m][2] = __float_as_uint(__ldg(Klo + o + 4));
      ah[m][3] = __float_as_uint(__ldg(Khi + o + 8 * (size_t)NR + 4)); al[m][3] = __float_as_uint(__ldg(Klo + o + 8 * (size_t)NR + 4));
    }
#pragma unroll
    for (int n = 0; n < NT; ++n) {
      // b0:(k = t, n = g) b1:(k = t + 4, n = g)
      const float x0 = X[(k0 + t) * BTP + n * 8 + g], x1 = X[(k0 + t + 4) * BTP + n * 8 + g];
      bh[n][0] = f2tf32(x0); bl[n][0] = f2tf32(x0 - __uint_as_float(bh[n][0]));
      bh[n][1] = f2tf32(x1); bl[n][1] = f2tf32(x1 - __uint_as_float(bh[n][1]));
    }
#pragma unroll
    for (int m = 0; m < 2; ++m)
#pragma unroll
      for (int n = 0; n < NT; ++n) {
#define PL_MMA_TF32(A, B)                                                                                     \
  asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};" \
               : "+f"(acc[m][n][0]), "+f"(acc[m][n][1]), "+f"(acc[m][n][2]), "+f"(acc[m][n][3])                \
               : "r"(A[m][0]), "r"(A[m][1]), "r"(A[m][2]), "r"(A[m][3]), "r"(B[n][0]), "r"(B[n][1]))
        PL_MMA_TF32(al, bh);       // small terms first
        PL_MMA_TF32(ah, bl);
        PL_MMA_TF32(ah, bh);
#undef PL_MMA_TF32
      }
  }
  // c0:(g, 2t) c1:(g, 2t+1) c2:(g+8, 2t) c3:(g+8, 2t+1)
#pragma unroll
  for (int m = 0; m < 2; ++m)
#pragma unroll
    for (int n = 0; n < NT; ++n) {
      float* y = Yb + (warp * 32 + m * 16 + g) * BTP + n * 8 + 2 * t;
      y[0] = acc[m][n][0]; y[1] = acc[m][n][1];
      y[8 * BTP] = acc[m][n][2]; y[8 * BTP + 1] = acc[m][n][3];
    }
}

// ------------------------------------------------------------------------------------------------ one stage
// y = derivative of row r at stage vector X (set-point P) + fr * fe[b]
template <typename T, int BT, int BTP, bool GEMM>
PL_DEV void stage_derivative(const Tables& tb, const T* vrow, int r, int4 w, bool own, const T* X, const T* P, T* Yb,
                             const T* kd_hi, const T* kd_lo, T fr, bool warp_force, const T* fe, T (&y)[BT]) {
  if constexpr (GEMM) {
    if (own) gemm_stage<BT, BTP>(kd_hi, kd_lo, tb.NR, X, Yb);
    __syncthreads();
    if (own) {
#pragma unroll
      for (int b = 0; b < BT; ++b) y[b] = Yb[r * BTP + b];
    }
  } else {
#pragma unroll
    for (int b = 0; b < BT; ++b) y[b] = T(0);
  }
  if (own) {
    row_terms<T, BT, BTP>(tb, vrow, r, w, X, P, y);
    if (warp_force) {
#pragma unroll
      for (int b = 0; b < BT; ++b) y[b] = fma(fr, fe[b], y[b]);
    }
  }
}

// ------------------------------------------------------------------------------------------------ forward kernel
template <typename T, int BT, int BTP, bool GEMM>
PL_DEV void forward_tile(const FwdArgs<T>& a, T* smem) {
  const Tables& tb = a.tb;
  const int NR = tb.NR;
  const int r = threadIdx.x;
  const int tile = blockIdx.x;
  const int inst = tile / a.tiles_per_inst;
  const int e0 = (tile - inst * a.tiles_per_inst) * BT;
  T* X0 = smem;
  T* X1 = X0 + NR * BTP;
  T* ering = X1 + NR * BTP;                      // [6][BT]
  const T* vrow = a.vals + (size_t)inst * tb.coef_stride + r;
  const int4 w = __ldg(tb.widths + (r >> 5));
  const T fr = __ldg(vrow + (size_t)(tb.WK + tb.WG + tb.WL + tb.WP) * NR);
  const bool warp_force = __any_sync(0xffffffffu, fr != T(0));
  const T h = a.h, h2 = a.h2, h6 = a.h6;
  const size_t tile_elems = (size_t)NR * BT;

  T q[BT], s[BT], y[BT];
  if (a.q_in) load_row<T, BT>(a.q_in + ((size_t)tile * NR + r) * BT, q);
  else {
    const T v = __ldg(a.ic + (size_t)inst * NR + r);
#pragma unroll
    for (int b = 0; b < BT; ++b) q[b] = v;
  }
#pragma unroll
  for (int b = 0; b < BT; ++b) X0[r * BTP + b] = q[b];

  // excitation ring: slot u % 6 holds sample u of the tile's BT excitations (u = t in full mode, 2t + half in half mode)
  const int ustep = a.half_mode ? 2 : 1;
  const bool loader = r < BT;
  const bool lvalid = loader && (e0 + r < a.E);
  const T* ecol = a.exc + (e0 + r);
  if (loader) {
    const long long u0 = (long long)ustep * a.t0;
    for (int k = 0; k <= ustep; ++k)
      ering[(int)((u0 + k) % 6) * BT + r] = lvalid ? __ldg(ecol + (u0 + k) * a.E) : T(0);
  }
  int save_cnt = -1, save_slot = 0;
  if (a.ckpt) {
    const int u = a.t0 + a.save_phase;
    save_slot = (u + a.save_every - 1) / a.save_every;
    save_cnt = save_slot * a.save_every - u;
  }
  int traj_cnt = -1, traj_slot = 0;
  if (a.user_traj) {
    traj_slot = (a.t0 + a.traj_every - 1) / a.traj_every;
    traj_cnt = traj_slot * a.traj_every - a.t0;
  }
  T* dense = a.dense ? a.dense + ((size_t)tile * NR + r) * BT : nullptr;
  const int* vm = a.varmap ? a.varmap + (size_t)inst * NR : nullptr;
  __syncthreads();

  for (int t = a.t0; t < a.t1; ++t) {
    const long long u = (long long)ustep * t;
    const T* fa = ering + (int)(u % 6) * BT;
    const T* fb = a.half_mode ? ering + (int)((u + 1) % 6) * BT : fa;
    const T* fc = ering + (int)((u + ustep) % 6) * BT;
    if (loader) {                                 // samples of step t + 1, visible after the next barrier
      for (int k = 1; k <= ustep; ++k) {
        const long long un = u + ustep + k;
        ering[(int)(un % 6) * BT + r] = lvalid ? __ldg(ecol + un * a.E) : T(0);
      }
    }
    stage_derivative<T, BT, BTP, GEMM>(tb, vrow, r, w, true, X0, X0, X1, a.kd_hi, a.kd_lo, fr, warp_force, fa, y);
#pragma unroll
    for (int b = 0; b < BT; ++b) { s[b] = y[b]; X1[r * BTP + b] = fma(h2, y[b], q[b]); }
    __syncthreads();
    stage_derivative<T, BT, BTP, GEMM>(tb, vrow, r, w, true, X1, X1, X0, a.kd_hi, a.kd_lo, fr, warp_force, fb, y);
#pragma unroll
    for (int b = 0; b < BT; ++b) { s[b] = fma(T(2), y[b], s[b]); X0[r * BTP + b] = fma(h2, y[b], q[b]); }
    __syncthreads();
    stage_derivative<T, BT, BTP, GEMM>(tb, vrow, r, w, true, X0, X0, X1, a.kd_hi, a.kd_lo, fr, warp_force, fb, y);
#pragma unroll
    for (int b = 0; b < BT; ++b) { s[b] = fma(T(2), y[b], s[b]); X1[r * BTP + b] = fma(h, y[b], q[b]); }
    __syncthreads();
    stage_derivative<T, BT, BTP, GEMM>(tb, vrow, r, w, true, X1, X1, X0, a.kd_hi, a.kd_lo, fr, warp_force, fc, y);
#pragma unroll
    for (int b = 0; b < BT; ++b) { q[b] = fma(h6, s[b] + y[b], q[b]); X0[r * BTP + b] = q[b]; }
    if (dense) {
      store_row<T, BT>(dense, q);
      dense += (size_t)a.n_tiles * tile_elems;
    }
    if (save_cnt == 0) {
      if (save_slot < a.save_slots) store_row<T, BT>(a.ckpt + ((size_t)save_slot * a.n_tiles + tile) * tile_elems + (size_t)r * BT, q);
      ++save_slot;
      save_cnt = a.save_every;
    }
    --save_cnt;
    if (traj_cnt == 0) {
      if (traj_slot < a.traj_slots) {
        const int gv = vm[r];
        if (gv >= 0) {
#pragma unroll
          for (int b = 0; b < BT; ++b)
            if (e0 + b < a.E) a.user_traj[((size_t)traj_slot * a.E + e0 + b) * a.n_glob + gv] = q[b];
        }
      }
      ++traj_slot;
      traj_cnt = a.traj_every;
    }
    --traj_cnt;
    __syncthreads();
  }
  if (a.user_final) {
    const int gv = vm[r];
    if (gv >= 0) {
#pragma unroll
      for (int b = 0; b < BT; ++b)
        if (e0 + b < a.E) a.user_final[(size_t)(e0 + b) * a.n_glob + gv] = q[b];
    }
  }
}

// ------------------------------------------------------------------------------------------------ adjoint kernel
// RK4 on the unforced adjoint ODE walking a recomputed chunk backwards (reference bwdStep, src/__GpuDriver.cu:746-783):
// forward state per stage = q_i, (q_i + q_{i-1})/2, q_{i-1}.  Threads [0, NR) own adjoint rows, threads [0, NPR) also
// own a forward row: they stream it from HBM one step ahead (register prefetch) into a 2-slot ring in shared memory
// and maintain the half-stage mean.  Literal-reference mode: the state with local index alias_local reads as the
// state after it, index -1 is `prev` (quirk A4).
template <typename T, int BT, int BTP, bool GEMM>
PL_DEV void backward_tile(const BwdArgs<T>& a, T* smem) {
  const Tables& tb = a.tb;
  const int NR = tb.NR, NPR = tb.NPR;
  const int r = threadIdx.x;
  const bool own = r < NR, ownf = r < NPR;
  const int tile = blockIdx.x;
  const int inst = tile / a.tiles_per_inst;
  const int e0 = (tile - inst * a.tiles_per_inst) * BT;
  T* X0 = smem;
  T* X1 = X0 + NR * BTP;
  T* R = X1 + NR * BTP;                         // [2][NPR][BTP] forward states i (even/odd): q_i is read in stage 1 only,
  T* H = R + 2 * NPR * BTP;                     //   so q_{i-2} replaces it in place; [NPR][BTP] half-stage mean
  const T* vrow = a.vals + (size_t)inst * tb.coef_stride + (own ? r : 0);
  const int4 w = own ? __ldg(tb.widths + (r >> 5)) : make_int4(0, 0, 0, 0);
  const T h = a.h, h2 = a.h2, h6 = a.h6;
  const size_t ftile = (size_t)NPR * BT;        // forward tile elements
  const size_t fstep = (size_t)a.n_tiles * ftile;
  const T* frow = a.dense + (size_t)tile * ftile + (size_t)r * BT;           // + i * fstep
  const T* prow = a.prev + (size_t)tile * ftile + (size_t)r * BT;
  auto slot = [&](int k) { return R + ((k + 2) & 1) * NPR * BTP; };

  T q[BT], s[BT], y[BT], pn[BT];
  if (own) {
    if (a.from_ic) {
      const T v = __ldg(a.ic + (size_t)inst * NR + r);
#pragma unroll
      for (int b = 0; b < BT; ++b) q[b] = v;
    } else {
      const T* lp = a.lam + ((size_t)tile * NR + r) * BT;      // plain loads: this kernel writes lam back
#pragma unroll
      for (int b = 0; b < BT; ++b) q[b] = lp[b];
    }
#pragma unroll
    for (int b = 0; b < BT; ++b) X0[r * BTP + b] = q[b];
  }
  const int top = a.len - 1;
  if (ownf && a.len > 0) {
    load_row<T, BT>(frow + (size_t)top * fstep, pn);
    T* ra = slot(top) + r * BTP;
#pragma unroll
    for (int b = 0; b < BT; ++b) ra[b] = pn[b];
    T prev_eff[BT];
#pragma unroll
    for (int b = 0; b < BT; ++b) prev_eff[b] = pn[b];
    if (top - 1 != a.alias_local) load_row<T, BT>(top >= 1 ? frow + (size_t)(top - 1) * fstep : prow, prev_eff);
    T* rb = slot(top - 1) + r * BTP;
#pragma unroll
    for (int b = 0; b < BT; ++b) { rb[b] = prev_eff[b]; H[r * BTP + b] = T(.5) * (pn[b] + prev_eff[b]); }
  }
  __syncthreads();

  for (int i = top; i >= 0; --i) {
    const bool more = i - 2 >= -1;              // state i-2 exists (index -1 = prev)
    if (ownf && more && (i - 2 != a.alias_local)) load_row<T, BT>(i >= 2 ? frow + (size_t)(i - 2) * fstep : prow, pn);
    const T* Pa = slot(i);
    const T* Pb = slot(i - 1);
    stage_derivative<T, BT, BTP, GEMM>(tb, vrow, r, w, own, X0, Pa, X1, a.kd_hi, a.kd_lo, T(0), false, nullptr, y);
    if (own) {
#pragma unroll
      for (int b = 0; b < BT; ++b) { s[b] = y[b]; X1[r * BTP + b] = fma(h2, y[b], q[b]); }
    }
    __syncthreads();
    stage_derivative<T, BT, BTP, GEMM>(tb, vrow, r, w, own, X1, H, X0, a.kd_hi, a.kd_lo, T(0), false, nullptr, y);
    if (own) {
#pragma unroll
      for (int b = 0; b < BT; ++b) { s[b] = fma(T(2), y[b], s[b]); X0[r * BTP + b] = fma(h2, y[b], q[b]); }
    }
    __syncthreads();
    stage_derivative<T, BT, BTP, GEMM>(tb, vrow, r, w, own, X0, H, X1, a.kd_hi, a.kd_lo, T(0), false, nullptr, y);
    if (own) {
#pragma unroll
      for (int b = 0; b < BT; ++b) { s[b] = fma(T(2), y[b], s[b]); X1[r * BTP + b] = fma(h, y[b], q[b]); }
    }
    __syncthreads();
    stage_derivative<T, BT, BTP, GEMM>(tb, vrow, r, w, own, X1, Pb, X0, a.kd_hi, a.kd_lo, T(0), false, nullptr, y);
    if (own) {
#pragma unroll
      for (int b = 0; b < BT; ++b) { q[b] = fma(h6, s[b] + y[b], q[b]); X0[r * BTP + b] = q[b]; }
    }
    // the slot of state i (read in stage 1 only) and H are not read in stage 4: install state i-2 (effective) and the
    // next half-stage mean
    if (ownf && more) {
      const T* rb = Pb + r * BTP;
      T* rn = slot(i - 2) + r * BTP;
      const bool aliased = (i - 2 == a.alias_local);
#pragma unroll
      for (int b = 0; b < BT; ++b) {
        const T vb = rb[b];
        const T vn = aliased ? vb : pn[b];
        rn[b] = vn;
        H[r * BTP + b] = T(.5) * (vb + vn);
      }
    }
    __syncthreads();
  }

  if (own) {
    store_row<T, BT>(a.lam + ((size_t)tile * NR + r) * BT, q);
    if (a.user_out) {
      const int gv = a.varmap[(size_t)inst * NR + r];
      if (gv >= 0) {
#pragma unroll
        for (int b = 0; b < BT; ++b)
          if (e0 + b < a.E) a.user_out[(size_t)(e0 + b) * a.n_glob + gv] = q[b];
      }
    }
  }
}

template <typename T, int BT, int BTP, bool GEMM, int MAXT>
__global__ void __launch_bounds__(MAXT) large_forward_kernel(const FwdArgs<T> a) {
  extern __shared__ __align__(16) unsigned char pl_smem[];
  forward_tile<T, BT, BTP, GEMM>(a, reinterpret_cast<T*>(pl_smem));
}
template <typename T, int BT, int BTP, bool GEMM, int MAXT>
__global__ void __launch_bounds__(MAXT) large_backward_kernel(const BwdArgs<T> a) {
  extern __shared__ __align__(16) unsigned char pl_smem[];
  backward_tile<T, BT, BTP, GEMM>(a, reinterpret_cast<T*>(pl_smem));
}

}  // namespace pasif_large
