// stage_kernels.cuh -- dense-operator path: one RK4 STAGE of the whole batch = one batched GEMM + one fused update.
//
// When a large component has a DENSE linear operator K' shared by every trajectory (BASELINE config 4, "many systems
// share a dense B"), K'.X is 2 N^2 flops per trajectory-stage against O(N) for everything else.  Keeping the state of
// a tile on chip (large_kernels.cuh) then re-streams K' from L2 for every 8-16 trajectories (~2 flop/byte); instead
// the state lives in HBM, trajectory-major [p][NR] (K-contiguous for both GEMM operands), and every stage is
//     Y[p][m]  = sum_k Kd[m][k] * X[p][k]                 a real GEMM  [NR x NR] . [NR x P]   (tensor cores)
//     update   : y = Y + Gamma/Lambda/Psi row terms + F e ; S, next stage vector, Q             (HBM-bound, fused)
// N/8 (fp64) .. N/4 (fp32) flop per state byte, so the HBM round trip of the state costs ~10 % of the GEMM at N = 512.
// This is the reference's own per-stage structure (cusparseSpMM over the [N x E] operand, src/__GpuDriver.cu:795-807)
// with the SpMM replaced by a tensor-core GEMM and its ~16 element-wise / atomic micro-kernels by ONE update kernel.
//
//   fp64: dgemm_tn_dmma_kernel    mma.sync.m8n8k4.f64 (DMMA), cp.async 3-stage shared-memory pipeline
//   fp32: sgemm_tn_3xtf32_kernel  3xTF32 split (hi*hi + lo*hi + hi*lo, fp32 accumulate) -- see stage_tcgen05.cuh for
//         the tcgen05/TMEM/TMA version used on sm_100a
#pragma once

#include "large_kernels.cuh"

namespace pasif_stage {

using pasif_large::Tables;

// ------------------------------------------------------------------------------------------------ fused stage update
template <typename T>
struct UpdArgs {
  Tables tb;             // row programs; K slots (WK) only for the rows OUTSIDE the GEMM band; NPR = row stride of the set-point arrays
  int y_lo, y_hi;        // rows [y_lo, y_hi) take their linear term from the GEMM result Y (the band of dense rows)
  const T* vals;         // [n_inst][coef_stride]
  const T* Y;            // [P][NR] GEMM result K'.X
  const T* X;            // [P][NR] stage vector the derivative is evaluated at
  const T* Pa;           // [P][NPR] set-point vector (forward: == X)
  const T* Pb;           // adjoint half stage: P = (Pa + Pb) / 2 ; else nullptr
  T* Q;                  // [P][NR] state
  T* S;                  // [P][NR] running stage sum
  T* Xn;                 // [P][NR] next stage vector (stage 4: the new state, or nullptr when Q alone receives it)
  T* Xn_hi;              // fp32 only: tf32 split of Xn for the next 3xTF32 GEMM
  T* Xn_lo;
  T* Qsave;              // stage 4: second copy of the new state (a chunk-trajectory slot of the adjoint recompute), or nullptr
  const T* fe;           // excitation samples of this stage [E] (nullptr: unforced)
  int stage;             // 1..4
  int E, n_traj;         // n_traj: trajectories of THIS launch (a contiguous part of the batch)
  long long p_base;      // global index of the part's first trajectory (instance = (p_base + p) / E)
  T h, h2, h6;
};

// grid (ceil(n_traj / (U * UPD_TG)), NR / 32), block (32, UPD_TG): a warp = 32 consecutive rows of U trajectories
constexpr int UPD_TG = 8;

// Body of the update for the U trajectories of one thread.  STAGE and HALF (adjoint half stage: set point = mean of two
// forward states) are compile-time, so the stage arithmetic and the set-point fetch carry no branches.  UNI (chosen by
// the host: U divides E and the part's first trajectory): the U trajectories of a thread always belong to ONE instance
// -- a coefficient is then fetched once per term instead of once per trajectory, and rows past the batch need no clamp
// (they map to valid excitations of the same instance).  Offsets inside the thread's block of U rows are 32-bit; only the
// block's base pointers are 64-bit (ncu of the first version: 74 % of the issue slots busy, 2/3 of them 64-bit address
// arithmetic and four 64-bit divisions per thread; DRAM at 43 %).
template <typename T, int U, int STAGE, bool HALF, bool UNI, bool SELF>
__device__ __forceinline__ void stage_update_body(const UpdArgs<T>& a, const int r, const long long p0, const unsigned inst0, const unsigned e0) {
  const Tables& tb = a.tb;
  const int NR = tb.NR, NPR = tb.NPR;
  const int4 w = __ldg(tb.widths + (r >> 5));
  const T* __restrict__ X = a.X + p0 * NR;
  const T* __restrict__ Pa = a.Pa + p0 * NPR;
  const T* __restrict__ Pb = HALF ? a.Pb + p0 * NPR : nullptr;
  const size_t o0 = (size_t)p0 * NR + r;
  const T* __restrict__ Yr = a.Y + o0;
  T* __restrict__ Qr = a.Q + o0;
  T* __restrict__ Sr = a.S + o0;
  T y[U], q[U], sv[U];
  unsigned ib[U], eb[U];                    // instance / excitation of trajectory b (general path)
  const bool in_band = r >= a.y_lo && r < a.y_hi;
  // every HBM operand of the row is requested up front (Y, Q, S of U trajectories: 3 U independent loads in flight per
  // thread); the row programs below only gather from the L1/L2-resident stage vector.  Arrays are padded to a multiple
  // of 128 trajectories, so rows past n_traj are readable (they recompute the last trajectory's coefficients).
#pragma unroll
  for (int b = 0; b < U; ++b) {
    y[b] = in_band ? Yr[b * NR] : T(0);
    q[b] = Qr[b * NR];
    sv[b] = STAGE >= 2 ? Sr[b * NR] : T(0);
  }
  if constexpr (!UNI) {
    const unsigned E = (unsigned)a.E;
    const long long last = a.n_traj - 1 - p0;                      // >= 0
    unsigned i = inst0, e = e0;
#pragma unroll
    for (int b = 0; b < U; ++b) {
      ib[b] = i; eb[b] = e;
      if (b < last) { if (++e == E) { e = 0; ++i; } }              // lanes past the batch stay on the last trajectory
    }
  }
  const size_t cs = (size_t)tb.coef_stride;
  const T* __restrict__ v0 = a.vals + r;
  const T* __restrict__ vu = v0 + inst0 * cs;
  auto coef = [&](int b, int o) -> T { return UNI ? T(0) : __ldg(v0 + ib[b] * cs + o); };
  auto P = [&](int b, int c) -> T {
    const T v = Pa[b * NPR + c];
    if constexpr (HALF) return T(.5) * (v + Pb[b * NPR + c]);
    else return v;
  };
  // operand gather from the stage vector: fp64 forms the address as (X + c) + b * NR (one wide multiply-add per index, then
  // uniform strides; measured 72 -> 69 us), fp32 as X + (b * NR + c) (the form its measured-best schedule was built from)
  auto xg = [&](int b, int c) -> T {
    if constexpr (sizeof(T) == 8) return (X + c)[b * NR];
    else return X[b * NR + c];
  };
  // sparse rows (x' = v, probes ...): their few linear terms as a row program, in CSR order like the reference
  // (trip counts are 0-3: the loops stay rolled)
  int off = 0;
#pragma unroll 1
  for (int j = 0; j < w.x; ++j) {
    const int col = __ldg(tb.k_col + j * NR + r);
    const T cu = UNI ? __ldg(vu + off + j * NR) : T(0);
#pragma unroll
    for (int b = 0; b < U; ++b) {
      const T kv = UNI ? cu : coef(b, off + j * NR);
      y[b] = fma(kv, xg(b, col), y[b]);
    }
  }
  off += tb.WK * NR;
  // SELF (forward system: the set point is the stage vector itself) and the class's tables say a == b (== c) in every
  // entry -- on-site x^2 / x^3 terms: one gather serves all operands of the term (same value, same products: bit-identical).
  // fp64 only (measured: update 73.5 -> 68.6 us per stage of 16384 trajectories); in fp32 the kernel is bound by the
  // latency of its dependent requests, not by their number, and the extra branches cost 2-6 us.
  constexpr bool DEDUP = sizeof(T) == 8;
  const int fl = DEDUP ? tb.flags : 0;
#pragma unroll 1
  for (int j = 0; j < w.y; ++j) {
    const int2 ix = __ldg(tb.g_idx + j * NR + r);
    const T cu = UNI ? __ldg(vu + off + j * NR) : T(0);
    if (DEDUP && SELF && (fl & 1)) {
#pragma unroll
      for (int b = 0; b < U; ++b) {
        const T gv = UNI ? cu : coef(b, off + j * NR);
        const T x = xg(b, ix.x);
        y[b] = fma(gv * x, x, y[b]);
      }
    } else {
#pragma unroll
      for (int b = 0; b < U; ++b) {
        const T gv = UNI ? cu : coef(b, off + j * NR);
        y[b] = fma(gv * xg(b, ix.x), P(b, ix.y), y[b]);
      }
    }
  }
  off += tb.WG * NR;
#pragma unroll 1
  for (int j = 0; j < w.z; ++j) {
    const int4 ix = __ldg(tb.l_idx + j * NR + r);
    const T cu = UNI ? __ldg(vu + off + j * NR) : T(0);
    if (DEDUP && SELF && (fl & 6) == 6) {
#pragma unroll
      for (int b = 0; b < U; ++b) {
        const T lv = UNI ? cu : coef(b, off + j * NR);
        const T x = xg(b, ix.x);
        y[b] = fma(lv * x * x, x, y[b]);
      }
    } else if (DEDUP && (fl & 4)) {
#pragma unroll
      for (int b = 0; b < U; ++b) {
        const T lv = UNI ? cu : coef(b, off + j * NR);
        const T pv = P(b, ix.y);
        y[b] = fma(lv * xg(b, ix.x) * pv, pv, y[b]);
      }
    } else {
#pragma unroll
      for (int b = 0; b < U; ++b) {
        const T lv = UNI ? cu : coef(b, off + j * NR);
        y[b] = fma(lv * xg(b, ix.x) * P(b, ix.y), P(b, ix.z), y[b]);
      }
    }
  }
  off += tb.WL * NR;
#pragma unroll 1
  for (int j = 0; j < w.w; ++j) {
    const int4 ix = __ldg(tb.p_idx + j * NR + r);
    const T cu = UNI ? __ldg(vu + off + j * NR) : T(0);
#pragma unroll
    for (int b = 0; b < U; ++b) {
      const T pv = UNI ? cu : coef(b, off + j * NR);
      y[b] = fma(pv * X[b * NR + ix.x] * P(b, ix.y) * P(b, ix.z), P(b, ix.w), y[b]);
    }
  }
  off += tb.WP * NR;
  if (a.fe) {
    const T cu = UNI ? __ldg(vu + off) : T(0);
#pragma unroll
    for (int b = 0; b < U; ++b) {
      const T fr = UNI ? cu : coef(b, off);
      y[b] = fma(fr, __ldg(a.fe + (UNI ? e0 + b : eb[b])), y[b]);
    }
  }
  T* __restrict__ Xn = a.Xn + o0;
#pragma unroll
  for (int b = 0; b < U; ++b) {
    T xn;
    if constexpr (STAGE == 1) { Sr[b * NR] = y[b]; xn = fma(a.h2, y[b], q[b]); }
    else if constexpr (STAGE == 2) { Sr[b * NR] = fma(T(2), y[b], sv[b]); xn = fma(a.h2, y[b], q[b]); }
    else if constexpr (STAGE == 3) { Sr[b * NR] = fma(T(2), y[b], sv[b]); xn = fma(a.h, y[b], q[b]); }
    else { xn = fma(a.h6, sv[b] + y[b], q[b]); Qr[b * NR] = xn; if (a.Qsave) a.Qsave[o0 + b * NR] = xn; }
    if (STAGE != 4 || a.Xn) Xn[b * NR] = xn;      // stage 4 may write the state array only (the host's stage_direct scheme)
    if constexpr (sizeof(T) == 4) {
      if (a.Xn_hi) {
        const float hi = __uint_as_float(pasif_large::f2tf32(xn));
        a.Xn_hi[o0 + b * NR] = hi;
        a.Xn_lo[o0 + b * NR] = __uint_as_float(pasif_large::f2tf32(xn - hi));
      }
    }
  }
}

// the kernel is HBM-bound: registers are capped at 64 so that 4 CTAs of 256 threads stay resident per SM (measured: fp32 +7 %, fp64 +3 %; a tighter cap spills)
template <typename T, int U, int STAGE, bool HALF, bool UNI, bool SELF>
__global__ void __launch_bounds__(32 * UPD_TG, 4) stage_update_kernel(const UpdArgs<T> a) {
  const int r = blockIdx.y * 32 + threadIdx.x;
  const long long p0 = ((long long)blockIdx.x * UPD_TG + threadIdx.y) * U;
  if (p0 >= a.n_traj) return;
  const unsigned pg0 = (unsigned)(a.p_base + p0);                  // the host refuses classes of 2^31 trajectories or more
  const unsigned inst0 = pg0 / (unsigned)a.E, e0 = pg0 - inst0 * (unsigned)a.E;
  stage_update_body<T, U, STAGE, HALF, UNI, SELF>(a, r, p0, inst0, e0);
}

// state initialisation: Q = X = ic (per instance), optional tf32 split
template <typename T>
__global__ void stage_init_kernel(const T* __restrict__ ic, T* Q, T* X, T* X_hi, T* X_lo, int NR, int E, long long n_traj,
                                  long long total, long long p_base) {
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= total) return;
  const long long p = idx / NR;
  const int r = (int)(idx - p * NR);
  const T v = p < n_traj ? ic[((p_base + p) / E) * NR + r] : T(0);
  if (Q) Q[idx] = v;
  X[idx] = v;
  if constexpr (sizeof(T) == 4) {
    if (X_hi) {
      const float hi = __uint_as_float(pasif_large::f2tf32(v));
      X_hi[idx] = hi;
      X_lo[idx] = __uint_as_float(pasif_large::f2tf32(v - hi));
    }
  }
}

// copy a state array into the stage-vector buffers (start of a recompute chunk): X = src (+ split)
template <typename T>
__global__ void stage_load_kernel(const T* __restrict__ src, T* Q, T* X, T* X_hi, T* X_lo, long long total) {
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= total) return;
  const T v = src[idx];
  if (Q) Q[idx] = v;
  X[idx] = v;
  if constexpr (sizeof(T) == 4) {
    if (X_hi) {
      const float hi = __uint_as_float(pasif_large::f2tf32(v));
      X_hi[idx] = hi;
      X_lo[idx] = __uint_as_float(pasif_large::f2tf32(v - hi));
    }
  }
}

// [p][NR] state -> user layout [slot][E][n_glob] through varmap [n_inst][NR]
template <typename T>
__global__ void stage_scatter_kernel(const T* __restrict__ Q, const int* __restrict__ varmap, T* __restrict__ out, int NR, int E,
                                     int n_glob, long long n_traj, long long p_base) {
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= n_traj * NR) return;
  const long long p = idx / NR;
  const int r = (int)(idx - p * NR);
  const long long inst = (p_base + p) / E;
  const int g = varmap[inst * NR + r];
  if (g >= 0) out[(p_base + p - inst * E) * n_glob + g] = Q[idx];
}

// ------------------------------------------------------------------------------------------------ fp64 GEMM (DMMA)
// Y[p][m] = sum_k A[m][k] * B[p][k];  A = Kd [Mpad][K] row-major, B = X [Ppad][K] row-major (both K-contiguous),
// Mpad % 64 == 0, Ppad % 128 == 0, K % 16 == 0.  CTA tile 64 (m) x 128 (p), 8 warps as 2 x 4, warp tile 32 x 32,
// 3-stage cp.async pipeline; shared-memory rows padded to 20 doubles so that every DMMA fragment load (8 rows x 4
// consecutive doubles per half-warp) is bank-conflict free.
// The tile height BM is a template parameter: 64 rows (8 warps as 2 x 4, three stages, two CTAs per SM) or 32 rows (4 warps,
// two stages, four CTAs per SM).  The C4 bands give 512 (forward) / 1024 (adjoint) tiles of 64 rows for 148 SMs -- 3.46 / 6.92
// tiles per SM, so the busiest SMs run 4 / 7 and the launch ends at 0.865 / 0.99 of a balanced one; with 32-row tiles the
// imbalance is at most one half-size tile.
constexpr int DG_BM = 64, DG_BN = 128, DG_BK = 16, DG_LD = 20, DG_STAGES = 3;
constexpr int DG_SMEM = DG_STAGES * (DG_BM + DG_BN) * DG_LD * 8;
template <int BM, int STAGES>
struct DgCfg {
  static constexpr int THREADS = BM * 4;                              // (BM / 32) x 4 warps
  static constexpr int SMEM = STAGES * (BM + DG_BN) * DG_LD * 8;
};

__device__ __forceinline__ void cp_async16(void* smem, const void* gmem) {
  const unsigned s = (unsigned)__cvta_generic_to_shared(smem);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(s), "l"(gmem));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N)); }

// lda / ldb: row strides of A / B (both pointers already offset to the first column of the K band)
template <int DG_BM = 64, int DG_STAGES = 3>
__global__ void __launch_bounds__(DG_BM * 4) dgemm_tn_dmma_kernel(const double* __restrict__ A, const double* __restrict__ B,
                                                                  double* __restrict__ Y, int M, int K, int ldy, int lda, int ldb) {
  constexpr int NTHR = DG_BM * 4;
  extern __shared__ __align__(16) unsigned char dg_smem[];
  double* sA = reinterpret_cast<double*>(dg_smem);                    // [STAGES][BM][LD]
  double* sB = sA + DG_STAGES * DG_BM * DG_LD;                        // [STAGES][BN][LD]
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int g = lane >> 2, t = lane & 3;
  const int wm = warp >> 2, wn = warp & 3;
  const int m0 = blockIdx.x * DG_BM;
  const long long p0 = (long long)blockIdx.y * DG_BN;
  const double* Ag = A + (size_t)m0 * lda;
  const double* Bg = B + (size_t)p0 * ldb;
  const int nkb = K / DG_BK;
  // programmatic dependent launch (no-ops for a normal launch): the successor may be scheduled, the predecessor has completed
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  asm volatile("griddepcontrol.wait;" ::: "memory");

  auto issue = [&](int kb, int st) {
    // A tile: 64 rows x 16 doubles = 512 16-byte chunks; B tile: 128 rows -> 1024 chunks
    double* a = sA + st * DG_BM * DG_LD;
    double* b = sB + st * DG_BN * DG_LD;
#pragma unroll
    for (int c = tid; c < DG_BM * 8; c += NTHR) {
      const int row = c >> 3, ch = c & 7;
      cp_async16(a + row * DG_LD + ch * 2, Ag + (size_t)row * lda + kb * DG_BK + ch * 2);
    }
#pragma unroll
    for (int c = tid; c < DG_BN * 8; c += NTHR) {
      const int row = c >> 3, ch = c & 7;
      cp_async16(b + row * DG_LD + ch * 2, Bg + (size_t)row * ldb + kb * DG_BK + ch * 2);
    }
  };

  double acc[4][4][2];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;

  for (int s = 0; s < DG_STAGES - 1; ++s) {
    if (s < nkb) issue(s, s);
    cp_async_commit();
  }
  for (int kb = 0; kb < nkb; ++kb) {
    cp_async_wait<DG_STAGES - 2>();
    __syncthreads();
    if (kb + DG_STAGES - 1 < nkb) issue(kb + DG_STAGES - 1, (kb + DG_STAGES - 1) % DG_STAGES);
    cp_async_commit();
    const double* a = sA + (kb % DG_STAGES) * DG_BM * DG_LD + (wm * 32 + g) * DG_LD + t;
    const double* b = sB + (kb % DG_STAGES) * DG_BN * DG_LD + (wn * 32 + g) * DG_LD + t;
#pragma unroll
    for (int kk = 0; kk < DG_BK; kk += 4) {
      double af[4], bf[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) af[i] = a[i * 8 * DG_LD + kk];
#pragma unroll
      for (int j = 0; j < 4; ++j) bf[j] = b[j * 8 * DG_LD + kk];
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j)
          asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                       : "+d"(acc[i][j][0]), "+d"(acc[i][j][1])
                       : "d"(af[i]), "d"(bf[j]));
    }
  }
  cp_async_wait<0>();
  // C fragment: row m = g, columns p = 2t, 2t+1
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int m = m0 + wm * 32 + i * 8 + g;
    if (m >= M) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const long long p = p0 + wn * 32 + j * 8 + 2 * t;
      Y[p * ldy + m] = acc[i][j][0];
      Y[(p + 1) * ldy + m] = acc[i][j][1];
    }
  }
}

// ------------------------------------------------------------------------------------------------ fp32 GEMM (3xTF32, mma.sync)
// Same tiling idea with m16n8k8 tf32: CTA tile 64 (m) x 128 (p), BK = 16 floats, operands pre-split into hi / lo
// (both exactly representable in tf32).  Kept as the portable reference for the tcgen05 kernel in stage_tcgen05.cuh.
constexpr int SG_BM = 64, SG_BN = 128, SG_BK = 16, SG_LD = 20, SG_STAGES = 3;
constexpr int SG_SMEM = SG_STAGES * 2 * (SG_BM + SG_BN) * SG_LD * 4;

__global__ void __launch_bounds__(256) sgemm_tn_3xtf32_kernel(const float* __restrict__ Ahi, const float* __restrict__ Alo,
                                                              const float* __restrict__ Bhi, const float* __restrict__ Blo,
                                                              float* __restrict__ Y, int M, int K, int ldy, int lda, int ldb) {
  extern __shared__ __align__(16) unsigned char dg_smem[];
  float* sAh = reinterpret_cast<float*>(dg_smem);                     // [STAGES][BM][LD]
  float* sAl = sAh + SG_STAGES * SG_BM * SG_LD;
  float* sBh = sAl + SG_STAGES * SG_BM * SG_LD;                       // [STAGES][BN][LD]
  float* sBl = sBh + SG_STAGES * SG_BN * SG_LD;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int g = lane >> 2, t = lane & 3;
  const int wm = warp >> 2, wn = warp & 3;                            // warp tile 32 (m) x 32 (p)
  const int m0 = blockIdx.x * SG_BM;
  const long long p0 = (long long)blockIdx.y * SG_BN;
  const int nkb = K / SG_BK;

  auto issue = [&](int kb, int st) {
    // rows of 16 floats = 4 chunks of 16 bytes
    for (int c = tid; c < SG_BM * 4; c += 256) {
      const int row = c >> 2, ch = c & 3;
      const size_t go = (size_t)(m0 + row) * lda + kb * SG_BK + ch * 4;
      cp_async16(sAh + (st * SG_BM + row) * SG_LD + ch * 4, Ahi + go);
      cp_async16(sAl + (st * SG_BM + row) * SG_LD + ch * 4, Alo + go);
    }
    for (int c = tid; c < SG_BN * 4; c += 256) {
      const int row = c >> 2, ch = c & 3;
      const size_t go = (size_t)(p0 + row) * ldb + kb * SG_BK + ch * 4;
      cp_async16(sBh + (st * SG_BN + row) * SG_LD + ch * 4, Bhi + go);
      cp_async16(sBl + (st * SG_BN + row) * SG_LD + ch * 4, Blo + go);
    }
  };

  float acc[2][4][4];
#pragma unroll
  for (int i = 0; i < 2; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j)
#pragma unroll
      for (int c = 0; c < 4; ++c) acc[i][j][c] = 0.f;

  for (int s = 0; s < SG_STAGES - 1; ++s) {
    if (s < nkb) issue(s, s);
    cp_async_commit();
  }
  for (int kb = 0; kb < nkb; ++kb) {
    cp_async_wait<SG_STAGES - 2>();
    __syncthreads();
    if (kb + SG_STAGES - 1 < nkb) issue(kb + SG_STAGES - 1, (kb + SG_STAGES - 1) % SG_STAGES);
    cp_async_commit();
    const int st = kb % SG_STAGES;
    const int ao = (st * SG_BM + wm * 32 + g) * SG_LD + t;
    const int bo = (st * SG_BN + wn * 32 + g) * SG_LD + t;
#pragma unroll
    for (int kk = 0; kk < SG_BK; kk += 8) {
      unsigned ah[2][4], al[2][4], bh[4][2], bl[4][2];
#pragma unroll
      for (int i = 0; i < 2; ++i) {
        const int o = ao + i * 16 * SG_LD + kk;
        ah[i][0] = __float_as_uint(sAh[o]);               al[i][0] = __float_as_uint(sAl[o]);
        ah[i][1] = __float_as_uint(sAh[o + 8 * SG_LD]);   al[i][1] = __float_as_uint(sAl[o + 8 * SG_LD]);
        ah[i][2] = __float_as_uint(sAh[o + 4]);           al[i][2] = __float_as_uint(sAl[o + 4]);
        ah[i][3] = __float_as_uint(sAh[o + 8 * SG_LD + 4]); al[i][3] = __float_as_uint(sAl[o + 8 * SG_LD + 4]);
      }
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int o = bo + j * 8 * SG_LD + kk;
        bh[j][0] = __float_as_uint(sBh[o]);     bl[j][0] = __float_as_uint(sBl[o]);
        bh[j][1] = __float_as_uint(sBh[o + 4]); bl[j][1] = __float_as_uint(sBl[o + 4]);
      }
#pragma unroll
      for (int i = 0; i < 2; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) {
#define PS_MMA(A_, B_)                                                                                                   \
  asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};" \
               : "+f"(acc[i][j][0]), "+f"(acc[i][j][1]), "+f"(acc[i][j][2]), "+f"(acc[i][j][3])                           \
               : "r"(A_[i][0]), "r"(A_[i][1]), "r"(A_[i][2]), "r"(A_[i][3]), "r"(B_[j][0]), "r"(B_[j][1]))
          PS_MMA(al, bh);
          PS_MMA(ah, bl);
          PS_MMA(ah, bh);
#undef PS_MMA
        }
    }
  }
  cp_async_wait<0>();
#pragma unroll
  for (int i = 0; i < 2; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const long long p = p0 + wn * 32 + j * 8 + 2 * t;
      const int m = m0 + wm * 32 + i * 16 + g;
      if (m < M) { Y[p * ldy + m] = acc[i][j][0]; Y[(p + 1) * ldy + m] = acc[i][j][1]; }
      if (m + 8 < M) { Y[p * ldy + m + 8] = acc[i][j][2]; Y[(p + 1) * ldy + m + 8] = acc[i][j][3]; }
    }
}

}  // namespace pasif_stage
