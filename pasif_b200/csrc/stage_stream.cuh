// stage_stream.cuh -- dense-operator path: the fused stage update as a PERSISTENT, bulk-copy-pipelined kernel.
//
// stage_update_kernel (stage_kernels.cuh) gives every thread one row of U trajectories and lets it fetch its operands
// itself.  ncu of that kernel (profiles/r02_ncu_c4d_stage_f32_bk16.csv): 350 warp instructions per 128 elements, most
// of them 64-bit address arithmetic, 5-7 long-scoreboard stalls per issue at 43 % occupancy, DRAM at 46 % -- the memory
// parallelism is tied to the occupancy and the operand gathers (index -> stage-vector element) are dependent global
// requests.  Here the memory traffic is decoupled from the threads:
//
//   * a CTA owns the rows of the system (thread r = row r, for the whole launch) and walks over TILES of TT consecutive
//     trajectories;
//   * every array of a tile is contiguous in HBM ([p][NR] layout: TT * NR values), so ONE thread (lane 0 of a producer
//     warp that owns no rows) moves a tile into shared memory with a handful of bulk copies (cp.async.bulk.shared.global
//     + mbarrier complete_tx), NBUF - 1 tiles ahead of the arithmetic -- tens of KB in flight per SM whatever the occupancy;
//   * the row programs gather from the tile in shared memory (any column of the trajectory's stage vector / set point is
//     there), so no dependent global request is left; results go out as plain coalesced stores (lane = row);
//   * because a thread keeps its row, the first term of every kind of its row program (index entries, and the
//     coefficients while the instance does not change) stays in registers from tile to tile;
//   * full[b] / empty[b] mbarriers per buffer: no CTA-wide barrier, the warps drift up to NBUF tiles apart.
//
// The arithmetic is the update kernel's, term by term in the same order: results are bit-identical
// (tests/test_large_gpu.py::test_stream_update_is_bit_identical).
#pragma once

#include "stage_kernels.cuh"
#include "stage_tcgen05.cuh"

namespace pasif_stage {

template <typename T>
struct StreamArgs {
  UpdArgs<T> u;
  int n_tiles;           // ceil(n_traj / TT)
  int nbuf;              // tiles in flight (shared-memory buffers), 2..STREAM_MAXBUF
  int yw;                // y_hi - y_lo: values of Y per trajectory that are loaded
  // offsets (in values of T) of the arrays inside one buffer; o_x == o_q when the state array is the stage vector
  // (stage 1 of the stage_direct scheme), o_pa == o_x for a forward system (the set point is the stage vector)
  int o_y, o_q, o_s, o_x, o_pa, o_pb, buf_elems;
  // dynamic tile scheduling (nullptr: tile = blockIdx.x + k * gridDim.x): sched[0] = next tile, sched[1] = CTAs that have
  // drawn their last tile; the last of them resets both for the next launch on the same stream.  A CTA that starts late
  // (its SM was busy with the GEMM of another part) then takes what is left instead of a fixed share of the batch.
  int* sched;
  // trajectory -> instance without a division: q = (n * e_mul) >> e_sh equals n / E for every n < 2^31 (e_mul = ceil(2^e_sh / E),
  // e_sh = 31 + ceil(log2 E); checked on the host); e_mul == 0: plain division
  unsigned e_mul;
  int e_sh;
};
constexpr int STREAM_MAXBUF = 4;

// wait of the 34 row warps of an SM for their tile: with a suspend-time hint the warp sleeps in the barrier unit instead of
// re-issuing the poll (ncu of the first version: 11 % of the kernel's warp instructions were try_wait / branch pairs)
__device__ __forceinline__ void mbar_wait_suspend(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "SWAIT_LOOP:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1, %2;\n"
      "@p bra SWAIT_DONE;\n"
      "bra SWAIT_LOOP;\n"
      "SWAIT_DONE:\n"
      "}\n" ::"r"(pasif_tc::smem_u32(bar)), "r"(parity), "r"(20000u)
      : "memory");
}

__device__ __forceinline__ void bulk_load(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(pasif_tc::smem_u32(dst)), "l"(src),
               "r"(bytes), "r"(pasif_tc::smem_u32(bar))
               : "memory");
}

// BLK: threads per CTA, at least NR + 32: warps [0, NR / 32) own the rows, the next warp is the producer.
// MINB: resident CTAs per SM the registers are capped for (up to 576 threads: 2 CTAs with 4-trajectory tiles, or 1 CTA with
// 8-trajectory tiles and twice the registers -- the same number of independent elements in flight per SM).
template <typename T, int TT, int BLK, int MINB, int STAGE, bool HALF, bool UNI>
__global__ void __launch_bounds__(BLK, MINB) stage_update_stream_kernel(const StreamArgs<T> a) {
  extern __shared__ __align__(128) unsigned char ss_raw[];
  uint64_t* full = reinterpret_cast<uint64_t*>(ss_raw);                 // [STREAM_MAXBUF]
  uint64_t* empty = full + STREAM_MAXBUF;                               // [STREAM_MAXBUF]
  int* tile_of = reinterpret_cast<int*>(empty + STREAM_MAXBUF);         // [STREAM_MAXBUF] tile in each buffer (-1: no more tiles)
  T* bufs = reinterpret_cast<T*>(ss_raw + 128);
  const UpdArgs<T>& u = a.u;
  const Tables& tb = u.tb;
  const int NR = tb.NR, NPR = tb.NPR;
  const int r = threadIdx.x;
  const int nbuf = a.nbuf;
  const int stride = gridDim.x;
  const int row_warps = NR >> 5;

  if (threadIdx.x == 0) {
    for (int s = 0; s < nbuf; ++s) { pasif_tc::mbar_init(&full[s], 1); pasif_tc::mbar_init(&empty[s], (uint32_t)row_warps); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  pasif_tc::pdl_launch_dependents();
  pasif_tc::pdl_wait();                                 // (programmatic dependent launch: the predecessor has completed from here on)

  if (r >= NR) {
    // ---- producer: lane 0 of the first warp without rows
    if (r != NR) return;
    const uint32_t row_bytes = (uint32_t)(TT * NR * sizeof(T)), p_bytes = (uint32_t)(TT * NPR * sizeof(T)), y_bytes = (uint32_t)(a.yw * sizeof(T));
    const bool ld_x = a.o_x != a.o_q, ld_pa = a.o_pa != a.o_x;
    uint32_t total = row_bytes + (uint32_t)TT * y_bytes;
    if (STAGE >= 2) total += row_bytes;
    if (ld_x) total += row_bytes;
    if (ld_pa) total += p_bytes;
    if (HALF) total += p_bytes;
    int buf = 0;
    uint32_t phase = 0;
    for (int tile = blockIdx.x;; tile += stride) {
      pasif_tc::mbar_wait(&empty[buf], phase ^ 1);                       // first pass over the buffers: free at once
      if (a.sched) tile = atomicAdd(a.sched, 1);
      if (tile >= a.n_tiles) {
        // end marker for the row warps; with dynamic scheduling the last CTA to get here re-arms the counters
        tile_of[buf] = -1;
        pasif_tc::mbar_arrive(&full[buf]);
        if (a.sched) {
          __threadfence();
          if (atomicAdd(a.sched + 1, 1) == (int)gridDim.x - 1) { a.sched[0] = 0; a.sched[1] = 0; __threadfence(); }
        }
        break;
      }
      tile_of[buf] = tile;
      const long long p0 = (long long)tile * TT;
      T* B = bufs + (size_t)buf * a.buf_elems;
      pasif_tc::mbar_expect_tx(&full[buf], total);
      bulk_load(B + a.o_q, u.Q + p0 * NR, row_bytes, &full[buf]);
      if (ld_x) bulk_load(B + a.o_x, u.X + p0 * NR, row_bytes, &full[buf]);
      if (STAGE >= 2) bulk_load(B + a.o_s, u.S + p0 * NR, row_bytes, &full[buf]);
      if (ld_pa) bulk_load(B + a.o_pa, u.Pa + p0 * NPR, p_bytes, &full[buf]);
      if (HALF) bulk_load(B + a.o_pb, u.Pb + p0 * NPR, p_bytes, &full[buf]);
      if (y_bytes) {
        // (rolled on purpose: one copy site with running pointers)
        const T* ysrc = u.Y + p0 * NR + u.y_lo;
        T* ydst = B + a.o_y;
#pragma unroll 1
        for (int t = 0; t < TT; ++t, ysrc += NR, ydst += a.yw) bulk_load(ydst, ysrc, y_bytes, &full[buf]);
      }
      if (++buf == nbuf) { buf = 0; phase ^= 1; }
    }
    return;
  }

  // ---- consumers: thread r owns row r
  const bool in_band = r >= u.y_lo && r < u.y_hi;
  const int4 w = __ldg(tb.widths + (r >> 5));
  const size_t cs = (size_t)tb.coef_stride;
  const unsigned E = (unsigned)u.E;
  const int off_g = tb.WK * NR, off_l = off_g + tb.WG * NR, off_p = off_l + tb.WL * NR, off_f = off_p + tb.WP * NR;
  // first entry of the K / Gamma / Lambda row programs of this row, 10 bits per index (NR, NPR <= 1024), so that they cost
  // two registers for the whole launch (zero when the warp has no such term: never used then)
  unsigned kg0 = 0, l0 = 0;
  if (w.x > 0) kg0 = (unsigned)__ldg(tb.k_col + r);
  if (w.y > 0) { const int2 g = __ldg(tb.g_idx + r); kg0 |= ((unsigned)g.x << 10) | ((unsigned)g.y << 20); }
  if (w.z > 0) { const int4 l = __ldg(tb.l_idx + r); l0 = (unsigned)l.x | ((unsigned)l.y << 10) | ((unsigned)l.z << 20); }
  const T* __restrict__ v0 = u.vals + r;
  // coefficients of those entries: fetched again only when the instance changes (UNI: one instance per tile)
  T ck = T(0), cg = T(0), cl = T(0), cf = T(0);
  unsigned cur_inst = 0xffffffffu;

  int buf = 0;
  uint32_t phase = 0;
  for (;;) {
    // the producer names the tile of a buffer before it arrives on (or hands the loads to) the buffer's barrier
    mbar_wait_suspend(&full[buf], phase);
    const int tile = tile_of[buf];
    if (tile < 0) break;
    const long long p0 = (long long)tile * TT;
    const int nb = (int)min((long long)TT, (long long)u.n_traj - p0);
    const unsigned pg0 = (unsigned)(u.p_base + p0);
    const unsigned inst0 = a.e_mul ? (unsigned)(((unsigned long long)pg0 * a.e_mul) >> a.e_sh) : pg0 / E;
    const unsigned e0 = pg0 - inst0 * E;
    unsigned ib[TT], eb[TT];
    if constexpr (UNI) {
      // one instance per tile (TT divides E and the part's offset): consecutive excitations of instance inst0; slots past the
      // batch cannot occur (the batch is a multiple of E, hence of TT)
#pragma unroll
      for (int b = 0; b < TT; ++b) { ib[b] = inst0; eb[b] = e0 + (unsigned)b; }
    } else {
      unsigned i = inst0, e = e0;
#pragma unroll
      for (int b = 0; b < TT; ++b) {
        ib[b] = i; eb[b] = e;
        if (b + 1 < nb) { if (++e == E) { e = 0; ++i; } }                // slots past the batch stay on the last trajectory
      }
    }
    const T* __restrict__ vu = v0 + inst0 * cs;
    if (UNI && inst0 != cur_inst) {
      cur_inst = inst0;
      if (w.x > 0) ck = __ldg(vu);
      if (w.y > 0) cg = __ldg(vu + off_g);
      if (w.z > 0) cl = __ldg(vu + off_l);
      if (u.fe) cf = __ldg(vu + off_f);
    }
    auto coef = [&](int b, int o) -> T { return __ldg(v0 + ib[b] * cs + o); };

    const T* __restrict__ B = bufs + (size_t)buf * a.buf_elems;
    const T* __restrict__ Xs = B + a.o_x;
    const T* __restrict__ Ps = B + a.o_pa;
    const T* __restrict__ Pbs = B + a.o_pb;
    auto P = [&](int b, unsigned c) -> T {
      const T v = Ps[b * NPR + c];
      if constexpr (HALF) return T(.5) * (v + Pbs[b * NPR + c]);
      else return v;
    };
    T y[TT], q[TT], sv[TT];
#pragma unroll
    for (int b = 0; b < TT; ++b) {
      y[b] = in_band ? B[a.o_y + b * a.yw + (r - u.y_lo)] : T(0);
      q[b] = B[a.o_q + b * NR + r];
      sv[b] = STAGE >= 2 ? B[a.o_s + b * NR + r] : T(0);
    }
    // ---- K terms outside the GEMM band (CSR order)
    if (w.x > 0) {
#pragma unroll
      for (int b = 0; b < TT; ++b) y[b] = fma(UNI ? ck : coef(b, 0), Xs[b * NR + (kg0 & 1023u)], y[b]);
#pragma unroll 1
      for (int j = 1; j < w.x; ++j) {
        const int col = __ldg(tb.k_col + j * NR + r);
        const T cu = UNI ? __ldg(vu + j * NR) : T(0);
#pragma unroll
        for (int b = 0; b < TT; ++b) y[b] = fma(UNI ? cu : coef(b, j * NR), Xs[b * NR + col], y[b]);
      }
    }
    // ---- Gamma
    if (w.y > 0) {
#pragma unroll
      for (int b = 0; b < TT; ++b) y[b] = fma((UNI ? cg : coef(b, off_g)) * Xs[b * NR + ((kg0 >> 10) & 1023u)], P(b, kg0 >> 20), y[b]);
#pragma unroll 1
      for (int j = 1; j < w.y; ++j) {
        const int2 ix = __ldg(tb.g_idx + j * NR + r);
        const T cu = UNI ? __ldg(vu + off_g + j * NR) : T(0);
#pragma unroll
        for (int b = 0; b < TT; ++b) y[b] = fma((UNI ? cu : coef(b, off_g + j * NR)) * Xs[b * NR + ix.x], P(b, ix.y), y[b]);
      }
    }
    // ---- Lambda
    if (w.z > 0) {
#pragma unroll
      for (int b = 0; b < TT; ++b)
        y[b] = fma((UNI ? cl : coef(b, off_l)) * Xs[b * NR + (l0 & 1023u)] * P(b, (l0 >> 10) & 1023u), P(b, l0 >> 20), y[b]);
#pragma unroll 1
      for (int j = 1; j < w.z; ++j) {
        const int4 ix = __ldg(tb.l_idx + j * NR + r);
        const T cu = UNI ? __ldg(vu + off_l + j * NR) : T(0);
#pragma unroll
        for (int b = 0; b < TT; ++b) y[b] = fma((UNI ? cu : coef(b, off_l + j * NR)) * Xs[b * NR + ix.x] * P(b, ix.y), P(b, ix.z), y[b]);
      }
    }
    // ---- Psi
#pragma unroll 1
    for (int j = 0; j < w.w; ++j) {
      const int4 ix = __ldg(tb.p_idx + j * NR + r);
      const T cu = UNI ? __ldg(vu + off_p + j * NR) : T(0);
#pragma unroll
      for (int b = 0; b < TT; ++b)
        y[b] = fma((UNI ? cu : coef(b, off_p + j * NR)) * Xs[b * NR + ix.x] * P(b, ix.y) * P(b, ix.z), P(b, ix.w), y[b]);
    }
    // ---- force
    if (u.fe) {
      if constexpr (UNI && TT == 4 && sizeof(T) == 4) {
        // the tile's four excitation samples are consecutive and 16-byte aligned (e0 and E are multiples of 4): one request
        const float4 f = __ldg(reinterpret_cast<const float4*>(u.fe + e0));
        y[0] = fma(cf, (T)f.x, y[0]); y[1] = fma(cf, (T)f.y, y[1]); y[2] = fma(cf, (T)f.z, y[2]); y[3] = fma(cf, (T)f.w, y[3]);
      } else if constexpr (UNI && TT == 2 && sizeof(T) == 8) {
        const double2 f = __ldg(reinterpret_cast<const double2*>(u.fe + e0));
        y[0] = fma(cf, (T)f.x, y[0]); y[1] = fma(cf, (T)f.y, y[1]);
      } else {
#pragma unroll
        for (int b = 0; b < TT; ++b) y[b] = fma(UNI ? cf : coef(b, off_f), __ldg(u.fe + eb[b]), y[b]);
      }
    }
    // the tile is in registers: hand the buffer back before the stores go out
    __syncwarp();
    if ((threadIdx.x & 31) == 0) pasif_tc::mbar_arrive(&empty[buf]);
    if (++buf == nbuf) { buf = 0; phase ^= 1; }

    T* __restrict__ Sg = u.S + (size_t)p0 * NR + r;
    T* __restrict__ Qg = u.Q + (size_t)p0 * NR + r;
    T* __restrict__ Xg = u.Xn ? u.Xn + (size_t)p0 * NR + r : nullptr;
    T* __restrict__ Vg = u.Qsave ? u.Qsave + (size_t)p0 * NR + r : nullptr;
#pragma unroll
    for (int b = 0; b < TT; ++b) {
      if (UNI || b < nb) {                                               // (UNI: every tile is full)
        T xn;
        if constexpr (STAGE == 1) { Sg[b * NR] = y[b]; xn = fma(u.h2, y[b], q[b]); }
        else if constexpr (STAGE == 2) { Sg[b * NR] = fma(T(2), y[b], sv[b]); xn = fma(u.h2, y[b], q[b]); }
        else if constexpr (STAGE == 3) { Sg[b * NR] = fma(T(2), y[b], sv[b]); xn = fma(u.h, y[b], q[b]); }
        else { xn = fma(u.h6, sv[b] + y[b], q[b]); Qg[b * NR] = xn; if (Vg) Vg[b * NR] = xn; }
        if (STAGE != 4 || Xg) Xg[b * NR] = xn;
      }
    }
  }
}

}  // namespace pasif_stage
