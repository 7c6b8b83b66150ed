// rk4_kernels.cuh -- register-resident RK4 forward / adjoint kernels for small polynomial ODE systems (sm_100a).
//
// One thread integrates ONE trajectory = (system instance, excitation) for the whole time range handed to the
// launch: state, RK4 stage vector and running stage sum live in registers, coefficients of the instance are
// loaded once into registers, and the only per-step memory traffic is one excitation sample (forward) or the
// forward state of the step (adjoint).  There are no atomics, no per-step launches and no inter-thread
// communication: the (instance x excitation) batch is embarrassingly parallel.
//
// The sparsity pattern of the system is a COMPILE-TIME traits class `P` (constexpr index tables), so every
// state/stage array index is a literal after unrolling and ptxas keeps the arrays in registers.  The same header
// is compiled ahead of time for the reference's example systems (patterns_aot.inc) and at run time through NVRTC
// for any other pattern (engine.cu); it must therefore stay free of #includes.
//
// What it replaces in the reference (per RK4 step: ~17 kernel launches + cuSPARSE SpMM + atomics):
//   fwdStep / derivatives / modterpolator      src/__GpuDriver.cu:693-713, 785-838, 840-890
//   SpTdV, applyForces, updateSlope, integrate src/kernels.cu:21-63, 71-88, 201-229
//   bwdStep + request/half/previous_setpoint   src/__GpuDriver.cu:746-783, src/kernels.cu:162-195
//
// Arithmetic contract (parity with the reference):
//   y_r  = sum_k K[r,k] x_k (CSR order) ; then Gamma, Lambda, Psi terms in hand-off order ; then F_r * e
//   each term  ((val * x_a) * p_b) [* p_c [* p_d]]  -- same association as SpTdV, last multiply fused into the add
//   mi   = q + dt * m                               (updateSlope)
//   q   += h6 * (m1 + 2 m2 + 2 m3 + m4)             (integrate; accumulated left to right in `s`, which is
//                                                    bit-identical because 2*m is exact)
#pragma once

#ifndef PASIF_BLOCK
#define PASIF_BLOCK 128
#endif

#define PASIF_DEV __device__ __forceinline__
// constexpr index table: after full unrolling the argument is a literal and the call folds to a constant
#define PASIF_TAB(name, ...)                       \
  PASIF_DEV static constexpr int name(int k) {     \
    constexpr int a[] = {__VA_ARGS__};             \
    return a[k];                                   \
  }

namespace pasif {

// ------------------------------------------------------------------------------------------------ launch arguments
template <typename T>
struct FwdArgs {
  const T* coef;        // [NC][n_inst]   instance-minor coefficient table (K, Gamma, Lambda, Psi, F values)
  const T* q_in;        // [N][n_traj]    start state (trajectory-minor SoA); nullptr -> broadcast `ic`
  const T* ic;          // [N][n_inst]
  T* q_out;             // [N][n_traj]    end state, may alias q_in; nullptr -> not written
  const T* exc;         // gated force samples: FULL-step mode [t][E] ; HALF-step mode [u][E], u = 2t + half
  int n_inst, E, n_traj;
  int t0, t1;           // steps t0 <= t < t1
  int half_mode;        // 0: stages use e[t], e[t], e[t], e[t+1] ; 1: f[2t], f[2t+1], f[2t+1], f[2t+2]
  T h, h2, h6;
  // periodic save after step t when (t + save_phase) % save_every == 0, into slot (t + save_phase) / save_every
  T* save_ptr;          // layout 0: [slot][N][n_traj]   layout 1 (user): [slot][E][n_glob] through varmap
  int save_every, save_phase, save_layout, save_slots;
  int n_glob;
  const int* varmap;    // [N][n_inst] global state index of local variable i of instance j
  // dense save (adjoint recompute): needed forward variables after EVERY step, [t - t0][n_traj][NV]
  // (one contiguous NV-vector per trajectory and step: a single 16-byte store / load when NV * sizeof(T) == 16)
  T* dense_ptr;
  // final state scattered to the user layout [e][n_glob] (getAmplitudes)
  T* final_user;
};

template <typename T>
struct BwdArgs {
  const T* coef;        // [NCA][n_inst]  adjoint coefficient table
  T* lam;               // [NA][n_traj]   adjoint state, in/out
  const T* dense;       // [i][n_traj][NV] forward state after step (t0 + i), i in [0, len)
  const T* prev;        // [NF][n_traj]   forward state standing in for index i = -1 (a checkpoint slot)
  int len;              // number of backward steps
  int alias_local;      // literal-reference mode: local index of the aliased scratch slot (quirk A3) or -1000000
  int n_inst, E, n_traj;
  T h, h2, h6;
};

// ------------------------------------------------------------------------------------------------ NV-vector access
template <typename T, int NV>
PASIF_DEV void store_vec(T* dst, const T (&v)[NV > 0 ? NV : 1]) {
  if constexpr (NV > 0 && (NV * sizeof(T)) % 16 == 0) {
    if constexpr (sizeof(T) == 4) {
#pragma unroll
      for (int k = 0; k < NV / 4; ++k) reinterpret_cast<float4*>(dst)[k] = make_float4(v[4 * k], v[4 * k + 1], v[4 * k + 2], v[4 * k + 3]);
    } else {
#pragma unroll
      for (int k = 0; k < NV / 2; ++k) reinterpret_cast<double2*>(dst)[k] = make_double2(v[2 * k], v[2 * k + 1]);
    }
  } else {
#pragma unroll
    for (int k = 0; k < NV; ++k) dst[k] = v[k];
  }
}
template <typename T, int NV>
PASIF_DEV void load_vec(const T* src, T (&v)[NV > 0 ? NV : 1]) {
  if constexpr (NV > 0 && (NV * sizeof(T)) % 16 == 0) {
    if constexpr (sizeof(T) == 4) {
#pragma unroll
      for (int k = 0; k < NV / 4; ++k) {
        const float4 t = __ldg(reinterpret_cast<const float4*>(src) + k);
        v[4 * k] = t.x; v[4 * k + 1] = t.y; v[4 * k + 2] = t.z; v[4 * k + 3] = t.w;
      }
    } else {
#pragma unroll
      for (int k = 0; k < NV / 2; ++k) {
        const double2 t = __ldg(reinterpret_cast<const double2*>(src) + k);
        v[2 * k] = t.x; v[2 * k + 1] = t.y;
      }
    }
  } else {
#pragma unroll
    for (int k = 0; k < NV; ++k) v[k] = __ldg(src + k);
  }
}

// ------------------------------------------------------------------------------------------------ packed fp32 pairs
// W = 2 (fp32 only): one thread integrates TWO trajectories -- excitations e, e+1 of the same instance -- as the two
// halves of 64-bit register pairs, on sm_100's packed FFMA2 / FMUL2 / FADD2 (PTX fma/mul/add.rn.f32x2).  Each half is
// rounded exactly like the scalar instruction, so results are bit-identical to W = 1; what changes is the issue
// rate: one instruction slot per two FMAs, and the coefficient (shared by the pair) enters as a scalar broadcast
// operand (`FFMA2 R0, R3.F32, R4.F32x2.HI_LO, R8.F32x2.HI_LO`), so it still costs one register.  The one-thread
// kernels are issue-bound (r01 ncu: 82 % of the issue slots for 75 % FMA-pipe activity); pairing halves the FP
// issue slots and shares loop / address arithmetic between two trajectories.
using ::fma;                                         // the scalar kernels keep the built-in fma next to the overloads below
struct f32x2 { float2 v; };
PASIF_DEV f32x2 fma(f32x2 a, f32x2 b, f32x2 c) { return f32x2{__ffma2_rn(a.v, b.v, c.v)}; }
PASIF_DEV f32x2 fma(float a, f32x2 b, f32x2 c) { return f32x2{__ffma2_rn(make_float2(a, a), b.v, c.v)}; }
PASIF_DEV f32x2 operator*(f32x2 a, f32x2 b) { return f32x2{__fmul2_rn(a.v, b.v)}; }
PASIF_DEV f32x2 operator*(float a, f32x2 b) { return f32x2{__fmul2_rn(make_float2(a, a), b.v)}; }
PASIF_DEV f32x2 operator+(f32x2 a, f32x2 b) { return f32x2{__fadd2_rn(a.v, b.v)}; }

template <typename T, int W> struct Vec { typedef T type; };
template <> struct Vec<float, 2> { typedef f32x2 type; };
template <typename V> PASIF_DEV V vzero() { return V(0); }
template <typename V, typename T> PASIF_DEV V bcast(T a) { return a; }
template <> PASIF_DEV f32x2 bcast<f32x2, float>(float a) { return f32x2{make_float2(a, a)}; }
template <> PASIF_DEV f32x2 vzero<f32x2>() { return f32x2{make_float2(0.f, 0.f)}; }
// W consecutive trajectories of a trajectory-minor array (8-byte access for a pair: n_traj and the pair index are even)
template <typename V, typename T> PASIF_DEV V ldv(const T* p) { return *reinterpret_cast<const V*>(p); }
template <typename V, typename T> PASIF_DEV void stv(T* p, V v) { *reinterpret_cast<V*>(p) = v; }
template <typename T> PASIF_DEV T half_of(T v, int) { return v; }
PASIF_DEV float half_of(f32x2 v, int w) { return w ? v.v.y : v.v.x; }

// ------------------------------------------------------------------------------------------------ lane-block classes
// P::G > 1: a component too large for one thread is integrated by a group of G consecutive threads (lane_split.hpp).
// Every lane owns P::N rows; before each derivative evaluation it publishes the P::NPUB slots other lanes read and
// fetches its P::H halo operands (slot pub_slot(hx_pub(h)) of lane (lane + hx_delta(h)) mod G) through a
// double-buffered shared-memory array: one group barrier per exchange.  Thread index p -> lane = p % G,
// trajectory = p / G = instance * E + e, virtual instance = instance * G + lane.
// (V: value type of a state slot -- T, or a packed pair in the two-trajectory kernels, which have no lane groups)
template <typename V>
struct LaneCtx {
  V* sx;            // [2][NPUB][PASIF_BLOCK]  state / adjoint-state exchange
  V* sp;            // [2][NPPUB][PASIF_BLOCK] set-point (forward state) exchange of the adjoint kernel
  int bx, bp;       // current buffer of each
  int tid, gbase, lane;
};
template <int G>
PASIF_DEV void group_sync(int tid) {
  if constexpr (G <= 32) __syncwarp();
  else if constexpr (G >= PASIF_BLOCK) __syncthreads();
  else asm volatile("bar.sync %0, %1;" ::"r"(1 + tid / G), "n"(G) : "memory");   // one named barrier per group of the CTA
}
template <typename V, class P, int NXE>
PASIF_DEV void exchange_x(V (&xe)[NXE], LaneCtx<V>& cx) {
  if constexpr (P::G > 1 && P::H > 0) {
    V* b = cx.sx + cx.bx * (P::NPUB * PASIF_BLOCK);
#pragma unroll
    for (int k = 0; k < P::NPUB; ++k) b[k * PASIF_BLOCK + cx.tid] = xe[P::pub_slot(k)];
    group_sync<P::G>(cx.tid);
#pragma unroll
    for (int h = 0; h < P::H; ++h)
      xe[P::N + h] = b[P::hx_pub(h) * PASIF_BLOCK + cx.gbase + ((cx.lane + P::hx_delta(h)) & (P::G - 1))];
    cx.bx ^= 1;
  }
}
template <typename V, class P, int NPE>
PASIF_DEV void exchange_p(V (&pe)[NPE], LaneCtx<V>& cx) {
  if constexpr (P::G > 1 && P::HP > 0) {
    V* b = cx.sp + cx.bp * (P::NPPUB * PASIF_BLOCK);
#pragma unroll
    for (int k = 0; k < P::NPPUB; ++k) b[k * PASIF_BLOCK + cx.tid] = pe[P::ppub_slot(k)];
    group_sync<P::G>(cx.tid);
#pragma unroll
    for (int h = 0; h < P::HP; ++h)
      pe[P::NV + h] = b[P::hp_pub(h) * PASIF_BLOCK + cx.gbase + ((cx.lane + P::hp_delta(h)) & (P::G - 1))];
    cx.bp ^= 1;
  }
}

// ------------------------------------------------------------------------------------------------ derivative
// y = K x + Gamma:(x,p) + Lambda:(x,p,p) + Psi:(x,p,p,p) + F fe        (forward: p == x)
// x / p may be longer than the row count: lane-block classes append the halo entries (operands owned by other lanes)
// V = T: one trajectory; V = f32x2: two trajectories of one instance (the coefficients c stay scalars)
template <typename V, typename T, class P, int NX, int NP_>
PASIF_DEV void derivative(const V (&x)[NX], const V (&p)[NP_], const T (&c)[P::NC > 0 ? P::NC : 1], V fe,
                          V (&y)[P::N]) {
#pragma unroll
  for (int r = 0; r < P::N; ++r) y[r] = vzero<V>();
#pragma unroll
  for (int k = 0; k < P::NNZK; ++k) {
    // unit coefficient (x' = v): fma(1, x, y) == y + x and fma(1, x, 0) == x, bit for bit (up to the sign of a zero) -- no
    // multiply, no coefficient register; the flags fold to constants after unrolling
    if (P::k_unit(k)) y[P::k_row(k)] = P::k_first(k) ? x[P::k_col(k)] : y[P::k_row(k)] + x[P::k_col(k)];
    else y[P::k_row(k)] = fma(c[P::OFF_K + k], x[P::k_col(k)], y[P::k_row(k)]);
  }
#pragma unroll
  for (int k = 0; k < P::NNZG; ++k)
    y[P::g_o(k)] = fma(c[P::OFF_G + k] * x[P::g_a(k)], p[P::g_b(k)], y[P::g_o(k)]);
#pragma unroll
  for (int k = 0; k < P::NNZL; ++k)
    y[P::l_o(k)] = fma(c[P::OFF_L + k] * x[P::l_a(k)] * p[P::l_b(k)], p[P::l_c(k)], y[P::l_o(k)]);
#pragma unroll
  for (int k = 0; k < P::NNZP; ++k)
    y[P::p_o(k)] = fma(c[P::OFF_P + k] * x[P::p_a(k)] * p[P::p_b(k)] * p[P::p_c(k)], p[P::p_d(k)], y[P::p_o(k)]);
#pragma unroll
  for (int k = 0; k < P::NF; ++k) y[P::f_row(k)] = fma(c[P::OFF_F + k], fe, y[P::f_row(k)]);
}

// one classical RK4 step of the forward system; fa/fb/fc = force sample of stage 1 / stages 2,3 / stage 4
template <typename V, typename T, class P>
PASIF_DEV void rk4_forward_step(V (&q)[P::N], const T (&c)[P::NC > 0 ? P::NC : 1], V fa, V fb, V fc, T h, T h2, T h6,
                                LaneCtx<V>& cx) {
  constexpr int NXE = P::N + P::H;
  V m[P::N], s[P::N], mi[NXE];
#pragma unroll
  for (int i = 0; i < P::N; ++i) mi[i] = q[i];
  exchange_x<V, P, NXE>(mi, cx);
  derivative<V, T, P, NXE, NXE>(mi, mi, c, fa, m);
#pragma unroll
  for (int i = 0; i < P::N; ++i) { s[i] = m[i]; mi[i] = fma(h2, m[i], q[i]); }
  exchange_x<V, P, NXE>(mi, cx);
  derivative<V, T, P, NXE, NXE>(mi, mi, c, fb, m);
#pragma unroll
  for (int i = 0; i < P::N; ++i) { s[i] = fma(T(2), m[i], s[i]); mi[i] = fma(h2, m[i], q[i]); }
  exchange_x<V, P, NXE>(mi, cx);
  derivative<V, T, P, NXE, NXE>(mi, mi, c, fb, m);
#pragma unroll
  for (int i = 0; i < P::N; ++i) { s[i] = fma(T(2), m[i], s[i]); mi[i] = fma(h, m[i], q[i]); }
  exchange_x<V, P, NXE>(mi, cx);
  derivative<V, T, P, NXE, NXE>(mi, mi, c, fc, m);
#pragma unroll
  for (int i = 0; i < P::N; ++i) q[i] = fma(h6, s[i] + m[i], q[i]);
}

// ------------------------------------------------------------------------------------------------ forward kernel
// PF: forward pattern.  DV: list of forward variables written by the dense save (DV::NV may be 0).
// W: trajectories per thread (2: packed fp32 pairs, E even; the dense save of a pair is ONE vector of 2 NV values in
// [variable][half] order -- written and read only by the pair kernels of the same class).  Lane-block classes (G > 1)
// pair the same lane of two consecutive trajectories; their halo exchange then moves 8-byte pairs.
template <typename T, class PF, class DV, int W = 1>
PASIF_DEV void forward_body(const FwdArgs<T>& a) {
  typedef typename Vec<T, W>::type V;
  const int tix = blockIdx.x * blockDim.x + threadIdx.x;
  const int p = tix * W;                           // first of the thread's W consecutive slots in the trajectory-minor arrays
  if (p >= a.n_traj) return;                       // lane-block classes: whole groups leave together (n_traj % G == 0)
  constexpr int G = PF::G;
  // thread -> (lane, trajectory): G consecutive threads are the lanes of one trajectory (W = 2: of the trajectory PAIR
  // 2j, 2j + 1, whose two values sit side by side in every array -- slot p = (j * G + lane) * 2 + half)
  const int lane = tix & (G - 1);
  const int traj = (tix / G) * W;
  const int e = traj % a.E;
  const int inst = (traj / a.E) * G + lane;       // virtual instance (G == 1: the instance itself)
  constexpr int N = PF::N;
  constexpr int NC = PF::NC > 0 ? PF::NC : 1;
  constexpr int SXN = (G > 1 && PF::H > 0) ? 2 * PF::NPUB * PASIF_BLOCK : 1;
  __shared__ V sm_x[SXN];
  LaneCtx<V> cx{sm_x, nullptr, 0, 0, (int)threadIdx.x, (int)threadIdx.x - lane, lane};

  T c[NC];
#pragma unroll
  for (int k = 0; k < PF::NC; ++k) c[k] = a.coef[(long long)k * a.n_inst + inst];
  V q[N];
  if (a.q_in) {
#pragma unroll
    for (int i = 0; i < N; ++i) q[i] = ldv<V>(a.q_in + (long long)i * a.n_traj + p);
  } else {
#pragma unroll
    for (int i = 0; i < N; ++i) q[i] = bcast<V>(a.ic[(long long)i * a.n_inst + inst]);
  }

  const T h = a.h, h2 = a.h2, h6 = a.h6;
  const long long E = a.E;
  // periodic-save bookkeeping without a per-step modulo
  int save_cnt = -1, save_slot = 0;
  if (a.save_ptr) {
    const int u = a.t0 + a.save_phase;
    save_slot = (u + a.save_every - 1) / a.save_every;          // first slot index with slot*every >= u
    save_cnt = save_slot * a.save_every - u;                    // steps (beyond t0) until that save fires
  }
  constexpr int NVD = DV::NV > 0 ? DV::NV : 1;
  constexpr int NVDW = DV::NV > 0 ? DV::NV * W : 1;
  T* dense = a.dense_ptr ? a.dense_ptr + (long long)p * DV::NV : nullptr;
  const long long dense_stride = (long long)a.n_traj * DV::NV;

  // everything that follows one RK4 step: dense save of the needed variables, periodic checkpoint / trajectory save
  auto after_step = [&]() {
    if (DV::NV > 0 && dense) {
      V dv[NVD];
#pragma unroll
      for (int v = 0; v < DV::NV; ++v) dv[v] = q[DV::nv(v)];
      store_vec<T, DV::NV * W>(dense, reinterpret_cast<const T(&)[NVDW]>(dv));
      dense += dense_stride;
    }
    if (save_cnt == 0) {
      if (save_slot < a.save_slots) {
        if (a.save_layout == 0) {
          T* dst = a.save_ptr + ((long long)save_slot * N) * a.n_traj + p;
#pragma unroll
          for (int i = 0; i < N; ++i) stv<V>(dst + (long long)i * a.n_traj, q[i]);
        } else {
#pragma unroll
          for (int w = 0; w < W; ++w) {
            T* dst = a.save_ptr + ((long long)save_slot * E + e + w) * a.n_glob;
#pragma unroll
            for (int i = 0; i < N; ++i) {
              const int g = a.varmap[(long long)i * a.n_inst + inst];
              if (G == 1 || g >= 0) dst[g] = half_of(q[i], w);                     // -1: padding slot of a lane
            }
          }
        }
      }
      ++save_slot;
      save_cnt = a.save_every;
    }
    --save_cnt;
  };

  if (!a.half_mode) {
    // gated excitation samples e[t .. t+4] in registers; the time loop is unrolled by 4 so that the next four samples
    // are requested a whole block (4 steps) before their first use and no register shuffling is left per step
    const T* ex = a.exc + (long long)a.t0 * E + e;
    V e0 = ldv<V>(ex), e1 = ldv<V>(ex + E), e2 = ldv<V>(ex + 2 * E), e3 = ldv<V>(ex + 3 * E), e4 = ldv<V>(ex + 4 * E);   // buffer is zero-padded past t1
    ex += 5 * E;
    int t = a.t0;
    // patterns with many terms (lane-block classes): one step per iteration, or the body outgrows the instruction cache
    // (ncu on the 256-DOF chain: 0.73 "no instruction" stalls per issue with the 4x body)
    if constexpr (PF::NC <= 40) {
      for (; t + 4 <= a.t1; t += 4) {
        const V n0 = ldv<V>(ex), n1 = ldv<V>(ex + E), n2 = ldv<V>(ex + 2 * E), n3 = ldv<V>(ex + 3 * E);           // e[t+5 .. t+8]
        ex += 4 * E;
        rk4_forward_step<V, T, PF>(q, c, e0, e0, e1, h, h2, h6, cx); after_step();
        rk4_forward_step<V, T, PF>(q, c, e1, e1, e2, h, h2, h6, cx); after_step();
        rk4_forward_step<V, T, PF>(q, c, e2, e2, e3, h, h2, h6, cx); after_step();
        rk4_forward_step<V, T, PF>(q, c, e3, e3, e4, h, h2, h6, cx); after_step();
        e0 = e4; e1 = n0; e2 = n1; e3 = n2; e4 = n3;
      }
    }
    for (; t < a.t1; ++t) {
      const V n0 = ldv<V>(ex);
      ex += E;
      rk4_forward_step<V, T, PF>(q, c, e0, e0, e1, h, h2, h6, cx); after_step();
      e0 = e1; e1 = e2; e2 = e3; e3 = e4; e4 = n0;
    }
  } else {
    // half-step mode (interpolation table): f[2t], f[2t+1], f[2t+2] -- preprocessed on the device by
    // prep_exc_half_kernel; two new samples per step, requested one step ahead
    const T* ex = a.exc + (long long)(2 * a.t0) * E + e;
    V f0 = ldv<V>(ex), f1 = ldv<V>(ex + E), f2 = ldv<V>(ex + 2 * E);
    ex += 3 * E;
    for (int t = a.t0; t < a.t1; ++t) {
      const V g1 = ldv<V>(ex), g2 = ldv<V>(ex + E);                                            // f[2t+3], f[2t+4]
      ex += 2 * E;
      rk4_forward_step<V, T, PF>(q, c, f0, f1, f2, h, h2, h6, cx); after_step();
      f0 = f2; f1 = g1; f2 = g2;
    }
  }

  if (a.q_out) {
#pragma unroll
    for (int i = 0; i < N; ++i) stv<V>(a.q_out + (long long)i * a.n_traj + p, q[i]);
  }
  if (a.final_user) {
#pragma unroll
    for (int w = 0; w < W; ++w) {
      T* dst = a.final_user + (long long)(e + w) * a.n_glob;
#pragma unroll
      for (int i = 0; i < N; ++i) {
        const int g = a.varmap[(long long)i * a.n_inst + inst];
        if (G == 1 || g >= 0) dst[g] = half_of(q[i], w);
      }
    }
  }
}

// ------------------------------------------------------------------------------------------------ adjoint kernel
// PA: adjoint pattern: N = adjoint state size, tensor "p" indices refer to the needed-forward-variable list
// (PA::NV entries, PA::nv(v) = local forward index).  RK4 on the unforced adjoint ODE, walking the recomputed
// chunk backwards; forward state per stage = q_i, (q_i + q_{i-1})/2, q_{i-1}  (src/__GpuDriver.cu:746-783).
template <typename T, class PA, int W = 1>
PASIF_DEV void backward_body(const BwdArgs<T>& a) {
  typedef typename Vec<T, W>::type V;
  const int tix = blockIdx.x * blockDim.x + threadIdx.x;
  const int p = tix * W;                           // slot of the thread in the trajectory-minor arrays (see forward_body)
  if (p >= a.n_traj) return;
  constexpr int G = PA::G;
  const int lane = tix & (G - 1);
  const int inst = ((tix / G) * W / a.E) * G + lane;   // virtual instance (G == 1: the instance itself)
  constexpr int N = PA::N;
  constexpr int NC = PA::NC > 0 ? PA::NC : 1;
  constexpr int NV = PA::NV > 0 ? PA::NV : 1;
  constexpr int NVW = PA::NV > 0 ? PA::NV * W : 1;
  constexpr int NXE = N + PA::H;                   // adjoint state + halo
  constexpr int NPE = NV + PA::HP;                 // needed forward variables + halo
  constexpr int SXN = (G > 1 && PA::H > 0) ? 2 * PA::NPUB * PASIF_BLOCK : 1;
  constexpr int SPN = (G > 1 && PA::HP > 0) ? 2 * PA::NPPUB * PASIF_BLOCK : 1;
  __shared__ V sm_x[SXN];
  __shared__ V sm_p[SPN];
  LaneCtx<V> cx{sm_x, sm_p, 0, 0, (int)threadIdx.x, (int)threadIdx.x - lane, lane};

  T c[NC];
#pragma unroll
  for (int k = 0; k < PA::NC; ++k) c[k] = a.coef[(long long)k * a.n_inst + inst];
  V lam[N];
#pragma unroll
  for (int i = 0; i < N; ++i) lam[i] = ldv<V>(a.lam + (long long)i * a.n_traj + p);

  const T h = a.h, h2 = a.h2, h6 = a.h6;
  const long long stride = (long long)PA::NV * a.n_traj;
  const T* d = a.dense + (long long)p * PA::NV + (long long)(a.len - 1) * stride;    // state index i = len-1
  V qa[NPE], qn[NV], qh[NPE];
  load_vec<T, PA::NV * W>(d, reinterpret_cast<T(&)[NVW]>(qa));
  exchange_p<V, PA, NPE>(qa, cx);
  // software prefetch: qn = state i-1 of the CURRENT iteration, loaded one iteration ahead
  if (a.len >= 2) load_vec<T, PA::NV * W>(d - stride, reinterpret_cast<T(&)[NVW]>(qn));
  else {
#pragma unroll
    for (int v = 0; v < PA::NV; ++v) qn[v] = ldv<V>(a.prev + (long long)PA::nv(v) * a.n_traj + p);
  }

  for (int i = a.len - 1; i >= 0; --i) {
    // qb = forward state of index i-1.  Literal-reference mode: when i-1 is the aliased scratch slot the reference
    // reads back what it has just written there, i.e. the current qa.
    V qb[NPE];
    const bool aliased = (i - 1 == a.alias_local);
#pragma unroll
    for (int v = 0; v < PA::NV; ++v) qb[v] = aliased ? qa[v] : qn[v];
    exchange_p<V, PA, NPE>(qb, cx);
    // prefetch state i-2 for the next iteration
    if (i >= 2) load_vec<T, PA::NV * W>(d - 2 * stride, reinterpret_cast<T(&)[NVW]>(qn));
    else if (i == 1) {
#pragma unroll
      for (int v = 0; v < PA::NV; ++v) qn[v] = ldv<V>(a.prev + (long long)PA::nv(v) * a.n_traj + p);
    }
    d -= stride;
#pragma unroll
    for (int v = 0; v < PA::NV + PA::HP; ++v) qh[v] = T(.5) * (qa[v] + qb[v]);

    V m[N], s[N], mi[NXE];
#pragma unroll
    for (int k = 0; k < N; ++k) mi[k] = lam[k];
    exchange_x<V, PA, NXE>(mi, cx);
    derivative<V, T, PA, NXE, NPE>(mi, qa, c, vzero<V>(), m);
#pragma unroll
    for (int k = 0; k < N; ++k) { s[k] = m[k]; mi[k] = fma(h2, m[k], lam[k]); }
    exchange_x<V, PA, NXE>(mi, cx);
    derivative<V, T, PA, NXE, NPE>(mi, qh, c, vzero<V>(), m);
#pragma unroll
    for (int k = 0; k < N; ++k) { s[k] = fma(T(2), m[k], s[k]); mi[k] = fma(h2, m[k], lam[k]); }
    exchange_x<V, PA, NXE>(mi, cx);
    derivative<V, T, PA, NXE, NPE>(mi, qh, c, vzero<V>(), m);
#pragma unroll
    for (int k = 0; k < N; ++k) { s[k] = fma(T(2), m[k], s[k]); mi[k] = fma(h, m[k], lam[k]); }
    exchange_x<V, PA, NXE>(mi, cx);
    derivative<V, T, PA, NXE, NPE>(mi, qb, c, vzero<V>(), m);
#pragma unroll
    for (int k = 0; k < N; ++k) lam[k] = fma(h6, s[k] + m[k], lam[k]);
#pragma unroll
    for (int v = 0; v < PA::NV + PA::HP; ++v) qa[v] = qb[v];
  }

#pragma unroll
  for (int i = 0; i < N; ++i) stv<V>(a.lam + (long long)i * a.n_traj + p, lam[i]);
}

struct NoDense {
  static constexpr int NV = 0, G = 1, H = 0, HP = 0, NPUB = 0, NPPUB = 0;
  PASIF_DEV static constexpr int nv(int) { return 0; }
};

}  // namespace pasif
