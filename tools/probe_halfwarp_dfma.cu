// probe_halfwarp_dfma.cu -- does a warp-wide DFMA with only 16 (8) active lanes occupy the fp64 pipe of its SM
// sub-partition for fewer cycles than a full warp?  One warp per CTA, 8 independent dependent chains per thread.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o build/probe_halfwarp_dfma tools/probe_halfwarp_dfma.cu
#include <cstdio>
#include <cuda_runtime.h>
template <typename T>
__global__ void k(T* out, int active, int iters, long long* cycles) {
  if ((int)threadIdx.x >= active) return;
  T a[8];
  for (int i = 0; i < 8; ++i) a[i] = (T)(threadIdx.x + i);
  const T b = (T)1.0000001, c = (T)1e-9;
  const long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int u = 0; u < 4; ++u)
#pragma unroll
      for (int i = 0; i < 8; ++i) a[i] = fma(a[i], b, c);
  }
  const long long t1 = clock64();
  T s = 0;
  for (int i = 0; i < 8; ++i) s += a[i];
  out[blockIdx.x * 32 + threadIdx.x] = s;
  if (threadIdx.x == 0 && blockIdx.x == 0) *cycles = t1 - t0;
}
int main() {
  double* out; long long* cyc;
  cudaMalloc(&out, 1024 * 32 * 8); cudaMalloc(&cyc, 8);
  const int iters = 4096;
  for (int blocks : {1, 148, 592}) {
    for (int active : {32, 16, 8, 4}) {
      long long h;
      k<double><<<blocks, 32>>>(out, active, iters, cyc); cudaDeviceSynchronize();
      k<double><<<blocks, 32>>>(out, active, iters, cyc); cudaDeviceSynchronize();
      cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
      printf("fp64 blocks %4d active lanes %2d: %.3f cycles per warp DFMA\n", blocks, active, (double)h / (iters * 32.0));
    }
  }
  for (int active : {32, 16}) {
    long long h;
    k<float><<<148, 32>>>((float*)out, active, iters, cyc); cudaDeviceSynchronize();
    cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
    printf("fp32 blocks  148 active lanes %2d: %.3f cycles per warp FFMA\n", active, (double)h / (iters * 32.0));
  }
  return 0;
}
