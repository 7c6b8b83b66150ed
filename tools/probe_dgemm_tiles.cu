// probe_dgemm_tiles.cu -- fp64 stage GEMM (DMMA, stage_kernels.cuh) at the C4 shapes with 64-row and 32-row tiles.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o build/probe_dgemm_tiles tools/probe_dgemm_tiles.cu
#include <cstdio>
#include <cstdlib>
#include <vector>
#include "../pasif_b200/csrc/stage_kernels.cuh"

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e_), __LINE__); exit(1); } } while (0)

template <int BM, int ST>
float run(const double* kd, const double* x, double* y, int NR, int P, int rows, int cols, int reps) {
  typedef pasif_stage::DgCfg<BM, ST> Cfg;
  auto fn = pasif_stage::dgemm_tn_dmma_kernel<BM, ST>;
  CK(cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM));
  dim3 grid(rows / BM, P / 128);
  cudaEvent_t a, b;
  CK(cudaEventCreate(&a)); CK(cudaEventCreate(&b));
  for (int i = 0; i < 3; ++i) fn<<<grid, Cfg::THREADS, Cfg::SMEM>>>(kd, x, y, rows, cols, NR, NR, NR);
  CK(cudaDeviceSynchronize());
  CK(cudaEventRecord(a));
  for (int i = 0; i < reps; ++i) fn<<<grid, Cfg::THREADS, Cfg::SMEM>>>(kd, x, y, rows, cols, NR, NR, NR);
  CK(cudaEventRecord(b));
  CK(cudaDeviceSynchronize());
  float ms;
  CK(cudaEventElapsedTime(&ms, a, b));
  return ms * 1000.f / reps;
}

int main() {
  const int NR = 544, P = 16384, reps = 30;
  double *kd, *x, *y, *y2;
  CK(cudaMalloc(&kd, 640 * NR * 8)); CK(cudaMalloc(&x, (size_t)P * NR * 8)); CK(cudaMalloc(&y, (size_t)P * NR * 8)); CK(cudaMalloc(&y2, (size_t)P * NR * 8));
  std::vector<double> h((size_t)P * NR);
  for (size_t i = 0; i < h.size(); ++i) h[i] = (double)((i * 2654435761u) % 1000) / 1000.0 - 0.5;
  CK(cudaMemcpy(x, h.data(), h.size() * 8, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(kd, h.data(), 640 * NR * 8, cudaMemcpyHostToDevice));
  struct { const char* name; int rows, cols; } shapes[] = {{"forward band 256 x 512", 256, 512}, {"adjoint band 512 x 256", 512, 256}};
  for (auto& s : shapes) {
    const double gf = 2.0 * P * s.rows * s.cols * 1e-9;
    float t;
    t = run<64, 3>(kd, x, y, NR, P, s.rows, s.cols, reps); printf("%-26s 64-row tiles, 3 stages  %7.2f us  %5.1f TFLOP/s\n", s.name, t, gf / t * 1e-3 * 1e3 / 1e3 * 1e3);
    t = run<64, 2>(kd, x, y, NR, P, s.rows, s.cols, reps); printf("%-26s 64-row tiles, 2 stages  %7.2f us\n", s.name, t);
    t = run<32, 3>(kd, x, y2, NR, P, s.rows, s.cols, reps); printf("%-26s 32-row tiles, 3 stages  %7.2f us\n", s.name, t);
    t = run<32, 2>(kd, x, y2, NR, P, s.rows, s.cols, reps); printf("%-26s 32-row tiles, 2 stages  %7.2f us\n", s.name, t);
    std::vector<double> a((size_t)P * NR), b((size_t)P * NR);
    CK(cudaMemcpy(a.data(), y, a.size() * 8, cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(b.data(), y2, b.size() * 8, cudaMemcpyDeviceToHost));
    size_t bad = 0;
    for (long long p = 0; p < P; ++p) for (int m = 0; m < s.rows; ++m) bad += a[p * NR + m] != b[p * NR + m];
    printf("   32-row vs 64-row tiles: %zu differing values\n", bad);
  }
  return 0;
}
