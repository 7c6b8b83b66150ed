// stage_tcgen05.cuh -- fp32 stage GEMM of the dense-operator path on the 5th-generation tensor cores (sm_100a):
//     Y[p][m] = sum_k Kd[m][k] * X[p][k]        3xTF32:  Xhi.Kdhi + Xlo.Kdhi + Xhi.Kdlo,  fp32 accumulate in TMEM
// UMMA M = 128 trajectories (TMEM lanes), N = 192 rows of K' (TMEM columns), K-block = 32 floats = one 128-byte
// swizzle atom.  Operands are K-major in shared memory, staged by TMA (cp.async.bulk.tensor.2d, SWIZZLE_128B) through
// a 2-stage full/empty mbarrier ring; tcgen05.mma is issued by one elected thread; the accumulator is read back with
// tcgen05.ld (32 lanes x 32 columns per warp and instruction) and written to Y as whole 128-byte lines per thread.
//   warp 0 / lane 0 : TMA producer          warp 1 / lane 0 : MMA issuer          warps 0-3 : epilogue
// Descriptor encodings follow the CUTLASS sm100 headers (cute/arch/mma_sm100_desc.hpp): shared-memory descriptor
// {start>>4 [0,14), LBO>>4 [16,30), SBO>>4 [32,46), version=1 [46,48), layout SWIZZLE_128B=2 [61,64)}, instruction
// descriptor {c_format F32=1 [4,6), a/b_format TF32=2 [7,10)/[10,13), a/b K-major, N>>3 [17,23), M>>4 [24,29)}.
#pragma once

#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace pasif_tc {

#ifndef PASIF_TC_STAGES_128
#define PASIF_TC_STAGES_128 3
#endif
constexpr int TC_BM = 128;          // trajectories per tile (UMMA M)
constexpr int TC_BK = 32;           // floats per k-block (128 bytes)
constexpr int TC_A_BYTES = TC_BM * TC_BK * 4;          // 16 KB (one of hi / lo)
// BN = rows of K' per tile (UMMA N): 192 for a full operator (544 rows = 3 tiles), 128 when only a band of dense rows goes
// through the GEMM (256 rows = 2 tiles, no padding work); the narrower tile leaves room for a third pipeline stage
// BK = floats per k-block: 32 (one 128-byte swizzle atom per row) or 16 (64-byte atoms).  With 256-row tiles a 32-float
// stage is 96 KB -- two stages, and the tensor pipe then waits for TMA + hi/lo split of the NEXT stage (ncu r01: tensor pipe
// 48 % active, 2.1 us per k-block against 1.1 us of MMA work); 16-float stages are 48 KB -- four stages, so the load of
// k-block kb+2 and the split of kb+1 run under the MMAs of kb.
template <int BN, int BK = TC_BK>
struct TcCfg {
  static constexpr int A_BYTES = TC_BM * BK * 4;        // one of hi / lo
  static constexpr int B_BYTES = BN * BK * 4;
  static constexpr int STAGE_BYTES = 2 * A_BYTES + 2 * B_BYTES;
  static constexpr int STAGES = BK == 16 ? (BN <= 128 ? 6 : 4) : (BN <= 128 ? PASIF_TC_STAGES_128 : 2);
  static constexpr int SMEM = STAGES * STAGE_BYTES + 1024 /* alignment slack */ + 256 /* barriers */;
  static constexpr int TMEM_COLS = BN <= 128 ? 128 : 256;   // power of two >= BN
};

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "WAIT_LOOP:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra WAIT_DONE;\n"
      "bra WAIT_LOOP;\n"
      "WAIT_DONE:\n"
      "}\n" ::"r"(smem_u32(bar)), "r"(parity)
      : "memory");
}
// Programmatic dependent launch: a kernel launched with cudaLaunchAttributeProgrammaticStreamSerialization may start while
// its predecessor in the stream is still draining; pdl_wait() blocks until that predecessor has completed and its writes
// are visible (every global access of the kernel comes after it), pdl_launch_dependents() lets the successor's CTAs be
// scheduled as soon as resources free up (it waits in its own pdl_wait()).  Both are no-ops for a normal launch.
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

__device__ __forceinline__ void tma_load_2d(void* dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(smem_u32(dst)),
               "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
               : "memory");
}
__device__ __forceinline__ uint64_t umma_desc_sw128(uint32_t saddr) {
  // K-major, SWIZZLE_128B: rows of 128 bytes, 8-row swizzle atoms 1024 bytes apart (SBO), LBO unused (1)
  return (uint64_t)((saddr >> 4) & 0x3FFFu) | ((uint64_t)1 << 16) | ((uint64_t)(1024 >> 4) << 32) | ((uint64_t)1 << 46) | ((uint64_t)2 << 61);
}
__device__ __forceinline__ uint64_t umma_desc_sw64(uint32_t saddr) {
  // K-major, SWIZZLE_64B: rows of 64 bytes, 8-row swizzle atoms 512 bytes apart (SBO); layout type 4
  return (uint64_t)((saddr >> 4) & 0x3FFFu) | ((uint64_t)1 << 16) | ((uint64_t)(512 >> 4) << 32) | ((uint64_t)1 << 46) | ((uint64_t)4 << 61);
}
template <int BK>
__device__ __forceinline__ uint64_t umma_desc_k(uint32_t saddr) {
  return BK == 16 ? umma_desc_sw64(saddr) : umma_desc_sw128(saddr);
}
__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n"
      "}\n" ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}

// grid (m tiles, p tiles), 128 threads.  mapA*: X hi / lo [Ppad][K]; mapB*: Kd hi / lo [Mpad][K]; p_row0: first row of
// this launch's part inside the X arrays; Y is already offset to the part.  Rows [m_base, M) of K' are computed
// (m_base a multiple of 32: the band of dense rows; Y columns outside it are not touched) over the K columns
// [k_base, k_base + K) (multiples of 32: the band of columns those rows use).
template <int TC_BN>
__global__ void __launch_bounds__(128, 1) sgemm_tn_3xtf32_tcgen05_kernel(const __grid_constant__ CUtensorMap mapAh,
                                                                         const __grid_constant__ CUtensorMap mapAl,
                                                                         const __grid_constant__ CUtensorMap mapBh,
                                                                         const __grid_constant__ CUtensorMap mapBl, float* __restrict__ Y,
                                                                         int M, int K, int ldy, int p_row0, int m_base, int k_base) {
  constexpr int TC_STAGES = TcCfg<TC_BN>::STAGES, TC_B_BYTES = TcCfg<TC_BN>::B_BYTES, TC_STAGE_BYTES = TcCfg<TC_BN>::STAGE_BYTES,
                TC_TMEM_COLS = TcCfg<TC_BN>::TMEM_COLS;
  extern __shared__ unsigned char tc_smem_raw[];
  unsigned char* base = reinterpret_cast<unsigned char*>(((uintptr_t)tc_smem_raw + 1023) & ~(uintptr_t)1023);
  uint64_t* full = reinterpret_cast<uint64_t*>(base + TC_STAGES * TC_STAGE_BYTES);
  uint64_t* empty = full + TC_STAGES;
  uint64_t* done = empty + TC_STAGES;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(done + 1);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int m0 = m_base + blockIdx.x * TC_BN;
  const int p0 = blockIdx.y * TC_BM;
  const int nkb = K / TC_BK;

  if (threadIdx.x == 0) {
    for (int s = 0; s < TC_STAGES; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
    mbar_init(done, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&mapAh) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&mapAl) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&mapBh) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&mapBl) : "memory");
  }
  if (warp == 2) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(TC_TMEM_COLS) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_d = *tmem_slot;

  if (warp == 0 && lane == 0) {
    // ---- TMA producer
    for (int kb = 0; kb < nkb; ++kb) {
      const int s = kb % TC_STAGES;
      const uint32_t ph = (kb / TC_STAGES) & 1;
      mbar_wait(&empty[s], ph ^ 1);                       // first pass: succeeds immediately (preceding phase)
      unsigned char* st = base + s * TC_STAGE_BYTES;
      mbar_expect_tx(&full[s], TC_STAGE_BYTES);
      tma_load_2d(st, &mapAh, &full[s], k_base + kb * TC_BK, p_row0 + p0);
      tma_load_2d(st + TC_A_BYTES, &mapAl, &full[s], k_base + kb * TC_BK, p_row0 + p0);
      tma_load_2d(st + 2 * TC_A_BYTES, &mapBh, &full[s], k_base + kb * TC_BK, m0);
      tma_load_2d(st + 2 * TC_A_BYTES + TC_B_BYTES, &mapBl, &full[s], k_base + kb * TC_BK, m0);
    }
  } else if (warp == 1 && lane == 0) {
    // ---- MMA issuer
    const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(TC_BN >> 3) << 17) | ((uint32_t)(TC_BM >> 4) << 24);
    for (int kb = 0; kb < nkb; ++kb) {
      const int s = kb % TC_STAGES;
      const uint32_t ph = (kb / TC_STAGES) & 1;
      mbar_wait(&full[s], ph);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      const uint32_t sa = smem_u32(base + s * TC_STAGE_BYTES);
      const uint32_t a_hi = sa, a_lo = sa + TC_A_BYTES, b_hi = sa + 2 * TC_A_BYTES, b_lo = b_hi + TC_B_BYTES;
#pragma unroll
      for (int kk = 0; kk < TC_BK / 8; ++kk) {            // UMMA K = 8 tf32 = 32 bytes inside the swizzle atom
        const uint32_t o = kk * 32;
        umma_tf32(tmem_d, umma_desc_sw128(a_lo + o), umma_desc_sw128(b_hi + o), idesc, (kb | kk) != 0);
        umma_tf32(tmem_d, umma_desc_sw128(a_hi + o), umma_desc_sw128(b_lo + o), idesc, 1u);
        umma_tf32(tmem_d, umma_desc_sw128(a_hi + o), umma_desc_sw128(b_hi + o), idesc, 1u);
      }
      umma_commit(&empty[s]);                             // frees the stage when these MMAs have read it
    }
    umma_commit(done);                                    // accumulator complete
  }

  // ---- epilogue: all four warps, warp w owns TMEM lanes [32w, 32w + 32)
  __syncwarp();
  mbar_wait(done, 0);
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  float* yrow = Y + (size_t)(p0 + warp * 32 + lane) * ldy;
#pragma unroll 1
  for (int c = 0; c < TC_BN / 32; ++c) {
    uint32_t v[32];
    const uint32_t taddr = tmem_d + ((uint32_t)(warp * 32) << 16) + (uint32_t)(c * 32);
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, %17, %18, %19, %20, "
        "%21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]), "=r"(v[10]),
          "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]),
          "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]),
          "=r"(v[31])
        : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
    const int mc = m0 + c * 32;
#pragma unroll
    for (int j = 0; j < 32; j += 4) {
      if (mc + j < M)                                     // M (= ldy) is a multiple of 32
        *reinterpret_cast<float4*>(yrow + mc + j) =
            make_float4(__uint_as_float(v[j]), __uint_as_float(v[j + 1]), __uint_as_float(v[j + 2]), __uint_as_float(v[j + 3]));
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 2) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_d), "r"(TC_TMEM_COLS) : "memory");
}

// ------------------------------------------------------------------------------------------------ fused hi / lo split
// Same GEMM, but the trajectory operand arrives as ONE fp32 array: the tile is loaded raw by TMA and split into its tf32
// hi / lo parts in shared memory, in place, by the four warps that later run the epilogue (they are idle during the main
// loop).  The update kernel then writes one stage vector instead of three and the GEMM reads 4 instead of 8 bytes per
// element of X.  192 threads: warps 0-3 convert + epilogue, warp 4 lane 0 TMA producer, warp 5 lane 0 MMA issuer.
//   full[s]  TMA bytes landed            -> converters
//   conv[s]  4 converter-warp arrivals   -> MMA issuer   (fence.proxy.async: generic-proxy writes visible to the tensor core)
//   empty[s] tcgen05.commit              -> producer
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ float tf32_rna(float x) {
  uint32_t u;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(u) : "f"(x));
  return __uint_as_float(u);
}

// Epilogue of the fused-split kernels.  A thread holds 32 consecutive columns of ONE trajectory after tcgen05.ld; stored
// directly, a warp instruction would touch 32 different lines (measured: 6 of the kernel's 27 us, tools/probe_gemm_parts.cu).
// Instead every 32 x 32 block goes through shared memory (the pipeline stages are free once `done` has fired; ebuf: 2 x 4 KB
// per warp, 1024-byte aligned), 128-byte swizzled so that the float4 writes of a quarter warp fall into different bank
// groups, and leaves as ONE bulk tensor store of 4 KB per warp and chunk (cp.async.bulk.tensor through mapY: 32 x 32 boxes).
// NCHUNK chunks of 32 columns from column m0 of Y on; columns from M on (the end of the band, a multiple of 32) are left alone.
template <int NCHUNK, int PROBE>
__device__ __forceinline__ void epilogue_bulk_store(uint32_t tmem_d, int warp, int lane, unsigned char* ebuf, const CUtensorMap* mapY, float* Y, int ldy,
                                                    int m0, int M, int p0) {
  float* yrow = Y + (size_t)(p0 + warp * 32 + lane) * ldy;
  // tcgen05.ld of chunk c + 1 is in flight while chunk c goes through shared memory (two register sets, loop unrolled by two)
  auto tmem_ld = [&](uint32_t (&v)[32], int c) {
    const uint32_t taddr = tmem_d + ((uint32_t)(warp * 32) << 16) + (uint32_t)(c * 32);
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, %17, %18, %19, %20, "
        "%21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]), "=r"(v[10]),
          "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]),
          "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]),
          "=r"(v[31])
        : "r"(taddr));
  };
  auto put = [&](const uint32_t (&v)[32], int c) {
    const int mc = m0 + c * 32;
    if (mc >= M) return;                                    // (M, the end of the band, and the tile origin are multiples of 32)
    if constexpr ((PROBE & 16) != 0) {                      // (probe: the direct stores of the first version)
#pragma unroll
      for (int j = 0; j < 32; j += 4)
        *reinterpret_cast<float4*>(yrow + mc + j) =
            make_float4(__uint_as_float(v[j]), __uint_as_float(v[j + 1]), __uint_as_float(v[j + 2]), __uint_as_float(v[j + 3]));
    } else if constexpr ((PROBE & 8) == 0) {
      unsigned char* eb = ebuf + (c & 1) * 4096;
      if (c >= 2) {                                         // the bulk store that read this buffer two chunks ago has finished reading it
        if (lane == 0) asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
        __syncwarp();
      }
#pragma unroll
      for (int j = 0; j < 8; ++j)
        *reinterpret_cast<float4*>(eb + lane * 128 + ((j ^ (lane & 7)) << 4)) =
            make_float4(__uint_as_float(v[4 * j]), __uint_as_float(v[4 * j + 1]), __uint_as_float(v[4 * j + 2]), __uint_as_float(v[4 * j + 3]));
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      __syncwarp();
      if (lane == 0) {
        asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];" ::"l"(mapY), "r"(smem_u32(eb)), "r"(mc),
                     "r"(p0 + warp * 32)
                     : "memory");
        asm volatile("cp.async.bulk.commit_group;" ::: "memory");
      }
    }
  };
  uint32_t va[32], vb[32];
  tmem_ld(va, 0);
#pragma unroll 1
  for (int c = 0; c < NCHUNK; c += 2) {
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
    if (c + 1 < NCHUNK) tmem_ld(vb, c + 1);
    put(va, c);
    if (c + 1 < NCHUNK) {
      asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
      if (c + 2 < NCHUNK) tmem_ld(va, c + 2);
      put(vb, c + 1);
    }
  }
  // shared memory must outlive the reads of the bulk stores; their global writes complete with the grid
  if (lane == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
}

// PROBE (tools/probe_gemm_parts.cu only; 0 in the library): 1 no hi/lo arithmetic, 2 no MMAs, 4 no TMA loads, 8 no stores
template <int TC_BN, int BK = TC_BK, int PROBE = 0>
__global__ void __launch_bounds__(192, 1) sgemm_tn_3xtf32_tcgen05_split_kernel(const __grid_constant__ CUtensorMap mapA,
                                                                               const __grid_constant__ CUtensorMap mapBh,
                                                                               const __grid_constant__ CUtensorMap mapBl,
                                                                               const __grid_constant__ CUtensorMap mapY, float* __restrict__ Y,
                                                                               int M, int K, int ldy, int p_row0, int m_base, int k_base, int nstages) {
  typedef TcCfg<TC_BN, BK> Cfg;
  constexpr int TC_B_BYTES = Cfg::B_BYTES, TC_STAGE_BYTES = Cfg::STAGE_BYTES, TC_TMEM_COLS = Cfg::TMEM_COLS, TC_A_BYTES = Cfg::A_BYTES;
  const int TC_STAGES = nstages;                        // 2 .. Cfg::STAGES: fewer stages leave shared memory for a co-resident update CTA
  extern __shared__ unsigned char tc_smem_raw[];
  unsigned char* base = reinterpret_cast<unsigned char*>(((uintptr_t)tc_smem_raw + 1023) & ~(uintptr_t)1023);
  uint64_t* full = reinterpret_cast<uint64_t*>(base + TC_STAGES * TC_STAGE_BYTES);
  uint64_t* empty = full + TC_STAGES;
  uint64_t* conv = empty + TC_STAGES;
  uint64_t* done = conv + TC_STAGES;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(done + 1);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int m0 = m_base + blockIdx.x * TC_BN;
  const int p0 = blockIdx.y * TC_BM;
  const int nkb = K / BK;

  if (threadIdx.x == 0) {
    for (int s = 0; s < TC_STAGES; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); mbar_init(&conv[s], 4); }
    mbar_init(done, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&mapA) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&mapBh) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&mapBl) : "memory");
  }
  if (warp == 2) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(TC_TMEM_COLS) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_d = *tmem_slot;
  pdl_launch_dependents();
  pdl_wait();                                           // barriers, TMEM and tensor maps were set up under the predecessor's tail

  if (warp == 4) {
    if (lane == 0) {
      // ---- TMA producer: raw X tile + Kd hi / lo tiles
      int s = 0;
      uint32_t ph = 0;
      for (int kb = 0; kb < nkb; ++kb, (++s == TC_STAGES ? (s = 0, ph ^= 1u) : 0u)) {
        mbar_wait(&empty[s], ph ^ 1);
        unsigned char* st = base + s * TC_STAGE_BYTES;
        if constexpr (PROBE & 4) { mbar_arrive(&full[s]); continue; }
        mbar_expect_tx(&full[s], TC_A_BYTES + 2 * TC_B_BYTES);
        tma_load_2d(st, &mapA, &full[s], k_base + kb * BK, p_row0 + p0);
        tma_load_2d(st + 2 * TC_A_BYTES, &mapBh, &full[s], k_base + kb * BK, m0);
        tma_load_2d(st + 2 * TC_A_BYTES + TC_B_BYTES, &mapBl, &full[s], k_base + kb * BK, m0);
      }
    }
  } else if (warp == 5) {
    if (lane == 0) {
      // ---- MMA issuer
      const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(TC_BN >> 3) << 17) | ((uint32_t)(TC_BM >> 4) << 24);
      int s = 0;
      uint32_t ph = 0;
      for (int kb = 0; kb < nkb; ++kb, (++s == TC_STAGES ? (s = 0, ph ^= 1u) : 0u)) {
        mbar_wait(&conv[s], ph);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const uint32_t sa = smem_u32(base + s * TC_STAGE_BYTES);
        const uint32_t a_hi = sa, a_lo = sa + TC_A_BYTES, b_hi = sa + 2 * TC_A_BYTES, b_lo = b_hi + TC_B_BYTES;
#pragma unroll
        for (int kk = 0; kk < BK / 8; ++kk) {
          if constexpr (PROBE & 2) break;
          const uint32_t o = kk * 32;
          umma_tf32(tmem_d, umma_desc_k<BK>(a_lo + o), umma_desc_k<BK>(b_hi + o), idesc, (kb | kk) != 0);
          umma_tf32(tmem_d, umma_desc_k<BK>(a_hi + o), umma_desc_k<BK>(b_lo + o), idesc, 1u);
          umma_tf32(tmem_d, umma_desc_k<BK>(a_hi + o), umma_desc_k<BK>(b_hi + o), idesc, 1u);
        }
        umma_commit(&empty[s]);
      }
      umma_commit(done);
    }
  } else {
    // ---- converters (warps 0-3): element-wise, so the 128-byte swizzle of the tile is preserved
    int s = 0;
    uint32_t ph = 0;
    for (int kb = 0; kb < nkb; ++kb, (++s == TC_STAGES ? (s = 0, ph ^= 1u) : 0u)) {
      mbar_wait(&full[s], ph);
      float4* a = reinterpret_cast<float4*>(base + s * TC_STAGE_BYTES);
      float4* l = reinterpret_cast<float4*>(base + s * TC_STAGE_BYTES + TC_A_BYTES);
#pragma unroll
      for (int i = 0; i < TC_A_BYTES / 16 / 128; ++i) {
        if constexpr (PROBE & 1) break;
        const int o = threadIdx.x + i * 128;
        const float4 v = a[o];
        float4 h, r;
        h.x = tf32_rna(v.x); h.y = tf32_rna(v.y); h.z = tf32_rna(v.z); h.w = tf32_rna(v.w);
        r.x = tf32_rna(v.x - h.x); r.y = tf32_rna(v.y - h.y); r.z = tf32_rna(v.z - h.z); r.w = tf32_rna(v.w - h.w);
        a[o] = h;
        l[o] = r;
      }
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      __syncwarp();
      if (lane == 0) mbar_arrive(&conv[s]);                 // one arrival per converter warp
    }
    // ---- epilogue: warp w owns TMEM lanes [32w, 32w + 32) = 32 trajectories
    mbar_wait(done, 0);
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    epilogue_bulk_store<TC_BN / 32, PROBE>(tmem_d, warp, lane, base + warp * 8192, &mapY, Y, ldy, m0, M, p0);
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 2) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_d), "r"(TC_TMEM_COLS) : "memory");
}

// ------------------------------------------------------------------------------------------------ cluster multicast variant
// The fused-split kernel is bound by what each SM pulls through L2: 16 KB of X and 32 KB of K' per k-block for 3.1 MFLOP.
// Two CTAs that work on the same rows of K' (neighbouring trajectory tiles) form a cluster (1, 2, 1): each loads HALF of
// the K' tile (64 rows of hi and of lo) and multicasts it into both shared memories, so every SM pulls 32 KB per k-block.
// A stage may only be overwritten when BOTH CTAs have consumed it: tcgen05.commit arrives on the `empty` barrier of both
// CTAs (multicast commit), which therefore counts 2.
__device__ __forceinline__ void tma_load_2d_mc(void* dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1, uint16_t mask) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1, {%3, %4}], [%2], %5;" ::"r"(
          smem_u32(dst)),
      "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "h"(mask)
      : "memory");
}
__device__ __forceinline__ void umma_commit_mc(uint64_t* bar, uint16_t mask) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(smem_u32(bar)), "h"(mask)
               : "memory");
}
__device__ __forceinline__ void cluster_sync() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}

// mapBh / mapBl: boxes of TC_BN / 2 rows.  grid.y must be even.
template <int TC_BN>
__global__ void __cluster_dims__(1, 2, 1) __launch_bounds__(192, 1)
    sgemm_tn_3xtf32_tcgen05_split_mc_kernel(const __grid_constant__ CUtensorMap mapA, const __grid_constant__ CUtensorMap mapBh,
                                            const __grid_constant__ CUtensorMap mapBl, float* __restrict__ Y, int M, int K, int ldy, int p_row0,
                                            int m_base, int k_base) {
  constexpr int TC_STAGES = TcCfg<TC_BN>::STAGES, TC_B_BYTES = TcCfg<TC_BN>::B_BYTES, TC_STAGE_BYTES = TcCfg<TC_BN>::STAGE_BYTES,
                TC_TMEM_COLS = TcCfg<TC_BN>::TMEM_COLS;
  extern __shared__ unsigned char tc_smem_raw[];
  unsigned char* base = reinterpret_cast<unsigned char*>(((uintptr_t)tc_smem_raw + 1023) & ~(uintptr_t)1023);
  uint64_t* full = reinterpret_cast<uint64_t*>(base + TC_STAGES * TC_STAGE_BYTES);
  uint64_t* empty = full + TC_STAGES;
  uint64_t* conv = empty + TC_STAGES;
  uint64_t* done = conv + TC_STAGES;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(done + 1);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int m0 = m_base + blockIdx.x * TC_BN;
  const int p0 = blockIdx.y * TC_BM;
  const int nkb = K / TC_BK;
  uint32_t rank;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(rank));

  if (threadIdx.x == 0) {
    for (int s = 0; s < TC_STAGES; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 2); mbar_init(&conv[s], 128); }
    mbar_init(done, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&mapA) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&mapBh) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&mapBl) : "memory");
  }
  if (warp == 2) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(TC_TMEM_COLS) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  cluster_sync();                                  // both CTAs' barriers exist before anyone multicasts into them
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_d = *tmem_slot;

  if (warp == 4) {
    if (lane == 0) {
      const int half_rows = TC_BN / 2, half_bytes = TC_B_BYTES / 2;
      for (int kb = 0; kb < nkb; ++kb) {
        const int s = kb % TC_STAGES;
        const uint32_t ph = (kb / TC_STAGES) & 1;
        mbar_wait(&empty[s], ph ^ 1);              // both CTAs have consumed this stage
        unsigned char* st = base + s * TC_STAGE_BYTES;
        mbar_expect_tx(&full[s], TC_A_BYTES + 2 * TC_B_BYTES);
        tma_load_2d(st, &mapA, &full[s], k_base + kb * TC_BK, p_row0 + p0);
        tma_load_2d_mc(st + 2 * TC_A_BYTES + rank * half_bytes, &mapBh, &full[s], k_base + kb * TC_BK, m0 + (int)rank * half_rows, (uint16_t)3);
        tma_load_2d_mc(st + 2 * TC_A_BYTES + TC_B_BYTES + rank * half_bytes, &mapBl, &full[s], k_base + kb * TC_BK, m0 + (int)rank * half_rows,
                       (uint16_t)3);
      }
    }
  } else if (warp == 5) {
    if (lane == 0) {
      const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(TC_BN >> 3) << 17) | ((uint32_t)(TC_BM >> 4) << 24);
      for (int kb = 0; kb < nkb; ++kb) {
        const int s = kb % TC_STAGES;
        const uint32_t ph = (kb / TC_STAGES) & 1;
        mbar_wait(&conv[s], ph);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const uint32_t sa = smem_u32(base + s * TC_STAGE_BYTES);
        const uint32_t a_hi = sa, a_lo = sa + TC_A_BYTES, b_hi = sa + 2 * TC_A_BYTES, b_lo = b_hi + TC_B_BYTES;
#pragma unroll
        for (int kk = 0; kk < TC_BK / 8; ++kk) {
          const uint32_t o = kk * 32;
          umma_tf32(tmem_d, umma_desc_sw128(a_lo + o), umma_desc_sw128(b_hi + o), idesc, (kb | kk) != 0);
          umma_tf32(tmem_d, umma_desc_sw128(a_hi + o), umma_desc_sw128(b_lo + o), idesc, 1u);
          umma_tf32(tmem_d, umma_desc_sw128(a_hi + o), umma_desc_sw128(b_hi + o), idesc, 1u);
        }
        umma_commit_mc(&empty[s], (uint16_t)3);    // this CTA is done with the stage: tell both producers
      }
      umma_commit(done);
    }
  } else {
    for (int kb = 0; kb < nkb; ++kb) {
      const int s = kb % TC_STAGES;
      const uint32_t ph = (kb / TC_STAGES) & 1;
      mbar_wait(&full[s], ph);
      float4* a = reinterpret_cast<float4*>(base + s * TC_STAGE_BYTES);
      float4* l = reinterpret_cast<float4*>(base + s * TC_STAGE_BYTES + TC_A_BYTES);
#pragma unroll
      for (int i = 0; i < TC_A_BYTES / 16 / 128; ++i) {
        const int o = threadIdx.x + i * 128;
        const float4 v = a[o];
        float4 h, r;
        h.x = tf32_rna(v.x); h.y = tf32_rna(v.y); h.z = tf32_rna(v.z); h.w = tf32_rna(v.w);
        r.x = tf32_rna(v.x - h.x); r.y = tf32_rna(v.y - h.y); r.z = tf32_rna(v.z - h.z); r.w = tf32_rna(v.w - h.w);
        a[o] = h;
        l[o] = r;
      }
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      mbar_arrive(&conv[s]);
    }
    mbar_wait(done, 0);
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    float* yrow = Y + (size_t)(p0 + warp * 32 + lane) * ldy;
#pragma unroll 1
    for (int c = 0; c < TC_BN / 32; ++c) {
      uint32_t v[32];
      const uint32_t taddr = tmem_d + ((uint32_t)(warp * 32) << 16) + (uint32_t)(c * 32);
      asm volatile(
          "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, %17, %18, %19, %20, "
          "%21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
          : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]), "=r"(v[10]),
            "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]),
            "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]),
            "=r"(v[31])
          : "r"(taddr));
      asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
      const int mc = m0 + c * 32;
#pragma unroll
      for (int j = 0; j < 32; j += 4) {
        if (mc + j < M)
          *reinterpret_cast<float4*>(yrow + mc + j) =
              make_float4(__uint_as_float(v[j]), __uint_as_float(v[j + 1]), __uint_as_float(v[j + 2]), __uint_as_float(v[j + 3]));
      }
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  cluster_sync();                                  // the peer may still multicast into this CTA / arrive on its barriers
  if (warp == 2) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_d), "r"(TC_TMEM_COLS) : "memory");
}

// ------------------------------------------------------------------------------------------------ CTA pairs (cta_group::2)
// The fused-split kernel above streams 40 KB per k-block and SM through L2 (8 KB of X, 32 KB of K' hi / lo) for 0.57 us of
// MMA work; 128 CTAs together ask L2 for 5.4 TB/s -- the per-k-block time follows the bytes (0.6 us with 128-row tiles and
// 24 KB, 0.95 us with 256-row tiles and 40 KB), not the MMA work, and TMA multicast inside a cluster of two does not reduce
// what L2 delivers (requests for the same line are merged there anyway).  A CTA PAIR does: tcgen05.mma.cta_group::2
// multiplies M = 256 trajectories (128 per CTA) by N = 256 rows of K' of which every CTA holds only HALF in its shared
// memory -- 8 KB of X + 16 KB of K' per k-block and SM.
//   cluster (2, 1, 1): rank 0 = leader (issues every MMA), rank 1 = peer; both CTAs load, split and run their epilogue
//   per CTA and stage (16-float k-blocks, SWIZZLE_64B): X hi | X lo | NT x (K' hi half | K' lo half)
//   NT = 1: one 256-row tile of K' per pair (256 TMEM columns, 6 stages); NT = 2: two tiles that share the X tile (512
//   columns, 4 stages) -- the 512-row band of the adjoint operator in ONE wave of 64 pairs
//   full[s]   (own CTA)       TMA bytes of this CTA landed          -> own converters
//   conv[s]   (leader's)      8 converter warps (4 per CTA) done    -> MMA issuer   (remote arrive from the peer)
//   empty[s]  (both CTAs)     multicast tcgen05.commit              -> both producers
//   done      (both CTAs)     multicast tcgen05.commit              -> both epilogues
// Results are bit-identical to the one-CTA kernels: the same products accumulate in the same order.
__device__ __forceinline__ uint32_t mapa_u32(uint32_t addr, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(rank));
  return r;
}
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t cluster_addr) {
  // default semantics (release at CTA scope) like CUTLASS's umma_arrive_2x1SM_sm0: what the leader's MMAs read was made
  // visible to the async proxy by fence.proxy.async before this arrive; the .release.cluster form costs a MEMBAR.ALL.GPU per
  // arrive and its .acquire.cluster counterpart an L1 invalidation per wait (ncu: 60 % of the stall samples of the first version)
  asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
// bounded wait: a protocol error ends in a trap after ~2 s instead of a hung GPU
__device__ __forceinline__ void mbar_wait_cluster(uint64_t* bar, uint32_t parity) {
  const uint32_t a = smem_u32(bar);
  const long long t0 = clock64();
  for (;;) {
    uint32_t ok;
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
        "selp.u32 %0, 1, 0, p;\n"
        "}\n"
        : "=r"(ok)
        : "r"(a), "r"(parity)
        : "memory");
    if (ok) return;
    if (clock64() - t0 > 4000000000ll) asm volatile("trap;");
  }
}
__device__ __forceinline__ void umma_tf32_2cta(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  const uint32_t z = 0;
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::2.kind::tf32 [%0], %1, %2, %3, {%5, %5, %5, %5, %5, %5, %5, %5}, p;\n"
      "}\n" ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate), "r"(z)
      : "memory");
}
__device__ __forceinline__ void umma_commit_2cta(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(smem_u32(bar)),
               "h"((uint16_t)3)
               : "memory");
}

template <int NT>
struct Tc2Cfg {
  static constexpr int BK = 16;
  static constexpr int A_BYTES = TC_BM * BK * 4;            // 8 KB: one of X hi / lo
  static constexpr int BH_BYTES = 128 * BK * 4;             // 8 KB: this CTA's half (128 rows) of one of K' hi / lo
  static constexpr int STAGE_BYTES = 2 * A_BYTES + NT * 2 * BH_BYTES;
  static constexpr int STAGES = NT == 1 ? 6 : 4;
  static constexpr int SMEM = STAGES * STAGE_BYTES + 1024 + 256;
  static constexpr int TMEM_COLS = 256 * NT;
};

// grid (trajectory tiles of 128, row tiles of 256 * NT), grid.x even: the pair = two neighbouring trajectory tiles;
// mapB*: boxes of 128 rows x 16 floats
template <int NT>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(192, 1)
    sgemm_tn_3xtf32_tcgen05_pair_kernel(const __grid_constant__ CUtensorMap mapA, const __grid_constant__ CUtensorMap mapBh,
                                        const __grid_constant__ CUtensorMap mapBl, const __grid_constant__ CUtensorMap mapY, float* __restrict__ Y, int M,
                                        int K, int ldy, int p_row0, int m_base, int k_base) {
  typedef Tc2Cfg<NT> Cfg;
  constexpr int BK = Cfg::BK, STAGES = Cfg::STAGES, A_BYTES = Cfg::A_BYTES, BH_BYTES = Cfg::BH_BYTES, STAGE_BYTES = Cfg::STAGE_BYTES;
  extern __shared__ unsigned char tc_smem_raw[];
  unsigned char* base = reinterpret_cast<unsigned char*>(((uintptr_t)tc_smem_raw + 1023) & ~(uintptr_t)1023);
  uint64_t* full = reinterpret_cast<uint64_t*>(base + STAGES * STAGE_BYTES);
  uint64_t* empty = full + STAGES;
  uint64_t* conv = empty + STAGES;
  uint64_t* done = conv + STAGES;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(done + 1);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int m0 = m_base + blockIdx.y * (256 * NT);
  const int p0 = blockIdx.x * TC_BM;
  const int nkb = K / BK;
  uint32_t rank;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(rank));

  if (threadIdx.x == 0) {
    for (int s = 0; s < STAGES; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); mbar_init(&conv[s], 8); }
    mbar_init(done, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&mapA) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&mapBh) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&mapBl) : "memory");
  }
  if (warp == 2) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(Cfg::TMEM_COLS) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  cluster_sync();                                  // both CTAs' barriers exist before anyone arrives on them
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_d = *tmem_slot;
  pdl_launch_dependents();
  pdl_wait();

  if (warp == 4) {
    if (lane == 0) {
      // ---- TMA producer (both CTAs): own X tile + own half of the K' hi / lo tiles
      for (int kb = 0; kb < nkb; ++kb) {
        const int s = kb % STAGES;
        const uint32_t ph = (kb / STAGES) & 1;
        mbar_wait_cluster(&empty[s], ph ^ 1);
        unsigned char* st = base + s * STAGE_BYTES;
        mbar_expect_tx(&full[s], A_BYTES + NT * 2 * BH_BYTES);
        tma_load_2d(st, &mapA, &full[s], k_base + kb * BK, p_row0 + p0);
#pragma unroll
        for (int nt = 0; nt < NT; ++nt) {
          const int row = m0 + nt * 256 + (int)rank * 128;
          tma_load_2d(st + 2 * A_BYTES + nt * 2 * BH_BYTES, &mapBh, &full[s], k_base + kb * BK, row);
          tma_load_2d(st + 2 * A_BYTES + nt * 2 * BH_BYTES + BH_BYTES, &mapBl, &full[s], k_base + kb * BK, row);
        }
      }
    }
  } else if (warp == 5) {
    if (lane == 0 && rank == 0) {
      // ---- MMA issuer (leader only): M = 256 (both CTAs' X tiles), N = 256 (both CTAs' halves of the K' tile)
      const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(256 >> 3) << 17) | ((uint32_t)(256 >> 4) << 24);
      for (int kb = 0; kb < nkb; ++kb) {
        const int s = kb % STAGES;
        const uint32_t ph = (kb / STAGES) & 1;
        mbar_wait_cluster(&conv[s], ph);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const uint32_t sa = smem_u32(base + s * STAGE_BYTES);
        const uint32_t a_hi = sa, a_lo = sa + A_BYTES;
#pragma unroll
        for (int kk = 0; kk < BK / 8; ++kk) {
          const uint32_t o = kk * 32;
#pragma unroll
          for (int nt = 0; nt < NT; ++nt) {
            const uint32_t b_hi = sa + 2 * A_BYTES + nt * 2 * BH_BYTES, b_lo = b_hi + BH_BYTES;
            const uint32_t d = tmem_d + (uint32_t)(nt * 256);
            umma_tf32_2cta(d, umma_desc_sw64(a_lo + o), umma_desc_sw64(b_hi + o), idesc, (kb | kk) != 0);
            umma_tf32_2cta(d, umma_desc_sw64(a_hi + o), umma_desc_sw64(b_lo + o), idesc, 1u);
            umma_tf32_2cta(d, umma_desc_sw64(a_hi + o), umma_desc_sw64(b_hi + o), idesc, 1u);
          }
        }
        umma_commit_2cta(&empty[s]);               // both CTAs' stage s may be refilled when these MMAs have read it
      }
      umma_commit_2cta(done);
    }
  } else {
    // ---- converters (warps 0-3 of both CTAs)
    const uint32_t conv0 = mapa_u32(smem_u32(&conv[0]), 0);          // the leader's conv barriers in the cluster window
    for (int kb = 0; kb < nkb; ++kb) {
      const int s = kb % STAGES;
      const uint32_t ph = (kb / STAGES) & 1;
      mbar_wait_cluster(&full[s], ph);
      float4* a = reinterpret_cast<float4*>(base + s * STAGE_BYTES);
      float4* l = reinterpret_cast<float4*>(base + s * STAGE_BYTES + A_BYTES);
#pragma unroll
      for (int i = 0; i < A_BYTES / 16 / 128; ++i) {
        const int o = threadIdx.x + i * 128;
        const float4 v = a[o];
        float4 h, r;
        h.x = tf32_rna(v.x); h.y = tf32_rna(v.y); h.z = tf32_rna(v.z); h.w = tf32_rna(v.w);
        r.x = tf32_rna(v.x - h.x); r.y = tf32_rna(v.y - h.y); r.z = tf32_rna(v.z - h.z); r.w = tf32_rna(v.w - h.w);
        a[o] = h;
        l[o] = r;
      }
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      __syncwarp();
      if (lane == 0) mbar_arrive_cluster(conv0 + (uint32_t)s * 8u);
    }
    // ---- epilogue: warp w owns TMEM lanes [32w, 32w + 32) of its own CTA = its own 128 trajectories
    mbar_wait_cluster(done, 0);
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    epilogue_bulk_store<NT * 8, 0>(tmem_d, warp, lane, base + warp * 8192, &mapY, Y, ldy, m0, M, p0);
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  cluster_sync();                                  // the peer's accumulator is read, nobody arrives on this CTA's barriers any more
  if (warp == 2) asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_d), "r"(Cfg::TMEM_COLS) : "memory");
}

}  // namespace pasif_tc
