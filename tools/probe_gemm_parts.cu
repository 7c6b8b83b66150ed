// probe_gemm_parts.cu -- where does the time of the fused-split tcgen05 GEMM go?  Times the library's kernel
// (stage_tcgen05.cuh) at the C4 shape with parts of its pipeline switched off (PROBE template flags).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o /tmp/probe_gemm_parts tools/probe_gemm_parts.cu -lcuda
#include <cstdio>
#include <cstdlib>
#include <vector>
#include "../pasif_b200/csrc/stage_tcgen05.cuh"

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e_), __LINE__); exit(1); } } while (0)

static CUtensorMap make_map(const void* base, uint64_t rows, uint64_t K, uint32_t box_rows, uint32_t box_k) {
  CUtensorMap m;
  const cuuint64_t dims[2] = {K, rows};
  const cuuint64_t strides[1] = {K * sizeof(float)};
  const cuuint32_t box[2] = {box_k, box_rows};
  const cuuint32_t estr[2] = {1, 1};
  CUresult rc = cuTensorMapEncodeTiled(&m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<void*>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                                       box_k == 16 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                                       CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (rc != CUDA_SUCCESS) { printf("encode failed %d\n", (int)rc); exit(1); }
  return m;
}

template <int BN, int BK, int PROBE>
float run(const float* kh, const float* kl, const float* x, float* y, int NR, int P, int m_lo, int m_hi, int stages, int reps, int K = 0) {
  typedef pasif_tc::TcCfg<BN, BK> Cfg;
  auto fn = pasif_tc::sgemm_tn_3xtf32_tcgen05_split_kernel<BN, BK, PROBE>;
  CK(cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM));
  const int ns = stages ? stages : Cfg::STAGES;
  const size_t smem = (size_t)ns * Cfg::STAGE_BYTES + 1280;
  CUtensorMap mA = make_map(x, P, NR, 128, BK), mBh = make_map(kh, 640, NR, BN, BK), mBl = make_map(kl, 640, NR, BN, BK);
  CUtensorMap mY = make_map(y, P, NR, 32, 32);
  dim3 grid((m_hi - m_lo) / BN, P / 128);
  cudaEvent_t a, b;
  CK(cudaEventCreate(&a)); CK(cudaEventCreate(&b));
  for (int i = 0; i < 3; ++i) fn<<<grid, 192, smem>>>(mA, mBh, mBl, mY, y, m_hi, K ? K : NR, NR, 0, m_lo, 0, ns);
  CK(cudaDeviceSynchronize());
  CK(cudaEventRecord(a));
  for (int i = 0; i < reps; ++i) fn<<<grid, 192, smem>>>(mA, mBh, mBl, mY, y, m_hi, K ? K : NR, NR, 0, m_lo, 0, ns);
  CK(cudaEventRecord(b));
  CK(cudaDeviceSynchronize());
  float ms;
  CK(cudaEventElapsedTime(&ms, a, b));
  return ms * 1000.f / reps;
}

// fixed cost of a launch of this shape: nothing / barriers + TMEM allocation / + one TMA round trip
__global__ void __launch_bounds__(192, 1) k_empty(int* out) { if (out && threadIdx.x == 999) *out = 1; }
__global__ void __launch_bounds__(192, 1) k_prologue(const __grid_constant__ CUtensorMap mapA, int tma, int cols) {
  using namespace pasif_tc;
  extern __shared__ unsigned char raw[];
  unsigned char* base = reinterpret_cast<unsigned char*>(((uintptr_t)raw + 1023) & ~(uintptr_t)1023);
  uint64_t* full = reinterpret_cast<uint64_t*>(base + 65536);
  uint32_t* slot = reinterpret_cast<uint32_t*>(full + 4);
  const int warp = threadIdx.x >> 5;
  if (threadIdx.x == 0) {
    for (int s = 0; s < 4; ++s) mbar_init(&full[s], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&mapA) : "memory");
  }
  if (warp == 2 && cols) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(slot)), "r"(cols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  if (tma) {
    if (threadIdx.x == 128) {
      mbar_expect_tx(&full[0], 8192);
      tma_load_2d(base, &mapA, &full[0], 0, blockIdx.y * 128);
    }
    mbar_wait(&full[0], 0);
  }
  __syncthreads();
  if (warp == 2 && cols) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(*slot), "r"(cols) : "memory");
}
template <typename F>
float time_launches(F launch, int reps) {
  cudaEvent_t a, b;
  CK(cudaEventCreate(&a)); CK(cudaEventCreate(&b));
  for (int i = 0; i < 3; ++i) launch();
  CK(cudaDeviceSynchronize());
  CK(cudaEventRecord(a));
  for (int i = 0; i < reps; ++i) launch();
  CK(cudaEventRecord(b));
  CK(cudaDeviceSynchronize());
  float ms;
  CK(cudaEventElapsedTime(&ms, a, b));
  return ms * 1000.f / reps;
}

template <int NT>
float run_pair(const float* kh, const float* kl, const float* x, float* y, int NR, int P, int m_lo, int m_hi, int reps) {
  typedef pasif_tc::Tc2Cfg<NT> Cfg;
  auto fn = pasif_tc::sgemm_tn_3xtf32_tcgen05_pair_kernel<NT>;
  CK(cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM));
  CUtensorMap mA = make_map(x, P, NR, 128, 16), mBh = make_map(kh, 640, NR, 128, 16), mBl = make_map(kl, 640, NR, 128, 16);
  CUtensorMap mY = make_map(y, P, NR, 32, 32);
  dim3 grid(P / 128, (m_hi - m_lo) / (256 * NT));
  cudaEvent_t a, b;
  CK(cudaEventCreate(&a)); CK(cudaEventCreate(&b));
  for (int i = 0; i < 3; ++i) fn<<<grid, 192, Cfg::SMEM>>>(mA, mBh, mBl, mY, y, m_hi, NR, NR, 0, m_lo, 0);
  CK(cudaDeviceSynchronize());
  CK(cudaEventRecord(a));
  for (int i = 0; i < reps; ++i) fn<<<grid, 192, Cfg::SMEM>>>(mA, mBh, mBl, mY, y, m_hi, NR, NR, 0, m_lo, 0);
  CK(cudaEventRecord(b));
  CK(cudaDeviceSynchronize());
  float ms;
  CK(cudaEventElapsedTime(&ms, a, b));
  return ms * 1000.f / reps;
}

int main(int argc, char** argv) {
  const int NR = 544, P = argc > 1 ? atoi(argv[1]) : 16384, reps = 50;
  CK(cudaFree(0));
  float *kh, *kl, *x, *y;
  CK(cudaMalloc(&kh, 640 * NR * 4)); CK(cudaMalloc(&kl, 640 * NR * 4));
  CK(cudaMalloc(&x, (size_t)P * NR * 4)); CK(cudaMalloc(&y, (size_t)P * NR * 4));
  std::vector<float> h((size_t)P * NR);
  for (size_t i = 0; i < h.size(); ++i) h[i] = (float)((i * 2654435761u) % 1000) / 1000.f - 0.5f;
  CK(cudaMemcpy(x, h.data(), h.size() * 4, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(kh, h.data(), 640 * NR * 4, cudaMemcpyHostToDevice));
  CK(cudaMemset(kl, 0, 640 * NR * 4));
#define R(BN, BK, PR, ST, LO, HI, what) printf("%-64s %7.2f us\n", what, run<BN, BK, PR>(kh, kl, x, y, NR, P, LO, HI, ST, reps));
  printf("fused-split GEMM, %d trajectories x 544 columns\n", P);
  R(256, 16, 0, 0, 256, 512, "256-row tile, bk16, 4 stages (forward band, default)");
  R(256, 16, 0, 3, 256, 512, "  3 stages");
  R(256, 16, 0, 2, 256, 512, "  2 stages");
  R(256, 16, 1, 0, 256, 512, "  no hi/lo arithmetic");
  R(256, 16, 2, 0, 256, 512, "  no MMAs");
  R(256, 16, 4, 0, 256, 512, "  no TMA loads");
  R(256, 16, 8, 0, 256, 512, "  no stores");
  R(256, 16, 16, 0, 256, 512, "  direct stores from registers (first version)");
  R(256, 16, 3, 0, 256, 512, "  no hi/lo, no MMAs (TMA + barriers)");
  R(256, 16, 5, 0, 256, 512, "  no hi/lo, no TMA (MMAs + barriers)");
  R(256, 16, 6, 0, 256, 512, "  no MMAs, no TMA (converters + barriers)");
  R(256, 16, 7, 0, 256, 512, "  barriers only");
  R(256, 16, 15, 0, 256, 512, "  barriers only, no stores");
  printf("%-64s %7.2f us\n", "  K = 32 (two k-blocks: launch + prologue + epilogue)", run<256, 16, 0>(kh, kl, x, y, NR, P, 256, 512, 0, reps, 32));
  printf("%-64s %7.2f us\n", "  K = 32, no stores", run<256, 16, 8>(kh, kl, x, y, NR, P, 256, 512, 0, reps, 32));
  printf("%-64s %7.2f us\n", "  K = 272 (half the k-blocks)", run<256, 16, 0>(kh, kl, x, y, NR, P, 256, 512, 0, reps, 272));
  R(128, 16, 0, 0, 256, 512, "128-row tiles, bk16, 6 stages (256 CTAs)");
  R(128, 16, 2, 0, 256, 512, "  no MMAs");
  R(128, 16, 4, 0, 256, 512, "  no TMA loads");
  R(256, 32, 0, 0, 256, 512, "256-row tile, bk32, 2 stages");
  R(256, 16, 0, 0, 0, 512, "adjoint band: 512 rows, bk16 (2 row tiles per trajectory tile)");
  R(256, 16, 16, 0, 0, 512, "  direct stores from registers (first version)");
  {
    CUtensorMap mA = make_map(x, P, NR, 128, 16);
    CK(cudaFuncSetAttribute(k_empty, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    CK(cudaFuncSetAttribute(k_prologue, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    dim3 grid(1, P / 128);
    printf("%-64s %7.2f us\n", "empty kernel, same grid / block, no dynamic shared memory", time_launches([&] { k_empty<<<grid, 192>>>(nullptr); }, reps));
    printf("%-64s %7.2f us\n", "empty kernel, 198 KB dynamic shared memory", time_launches([&] { k_empty<<<grid, 192, 198 * 1024>>>(nullptr); }, reps));
    printf("%-64s %7.2f us\n", "  + barriers, tensor-map prefetch, __syncthreads", time_launches([&] { k_prologue<<<grid, 192, 198 * 1024>>>(mA, 0, 0); }, reps));
    printf("%-64s %7.2f us\n", "  + TMEM alloc / dealloc of 256 columns", time_launches([&] { k_prologue<<<grid, 192, 198 * 1024>>>(mA, 0, 256); }, reps));
    printf("%-64s %7.2f us\n", "  + one TMA load of the X tile and the wait for it", time_launches([&] { k_prologue<<<grid, 192, 198 * 1024>>>(mA, 1, 256); }, reps));
  }
  printf("%-64s %7.2f us\n", "CTA pairs (cta_group::2), forward band, one 256-row tile", run_pair<1>(kh, kl, x, y, NR, P, 256, 512, reps));
  printf("%-64s %7.2f us\n", "CTA pairs, adjoint band, two 256-row tiles per pair", run_pair<2>(kh, kl, x, y, NR, P, 0, 512, reps));
  printf("%-64s %7.2f us\n", "CTA pairs, adjoint band, one tile per pair (2 waves)", run_pair<1>(kh, kl, x, y, NR, P, 0, 512, reps));
  // the bulk-store epilogue writes what the direct stores wrote
  std::vector<float> ya((size_t)P * NR), yb((size_t)P * NR);
  CK(cudaMemset(y, 0, (size_t)P * NR * 4));
  run<256, 16, 16>(kh, kl, x, y, NR, P, 0, 512, 0, 1);
  CK(cudaMemcpy(ya.data(), y, ya.size() * 4, cudaMemcpyDeviceToHost));
  CK(cudaMemset(y, 0, (size_t)P * NR * 4));
  run<256, 16, 0>(kh, kl, x, y, NR, P, 0, 512, 0, 1);
  CK(cudaMemcpy(yb.data(), y, yb.size() * 4, cudaMemcpyDeviceToHost));
  size_t bad = 0, nz = 0;
  for (size_t i = 0; i < ya.size(); ++i) { bad += ya[i] != yb[i]; nz += ya[i] != 0.f; }
  {
    CK(cudaMemset(y, 0, (size_t)P * NR * 4));
    run_pair<2>(kh, kl, x, y, NR, P, 0, 512, 1);
    std::vector<float> yc((size_t)P * NR);
    CK(cudaMemcpy(yc.data(), y, yc.size() * 4, cudaMemcpyDeviceToHost));
    size_t badp = 0;
    for (size_t i = 0; i < ya.size(); ++i) badp += ya[i] != yc[i];
    printf("CTA-pair kernel vs direct stores: %zu differing values\n", badp);
  }
  printf("bulk-store epilogue vs direct stores: %zu differing values of %zu (%zu non-zero)\n", bad, ya.size(), nz);
  return 0;
}
