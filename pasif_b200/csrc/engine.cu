// engine.cu -- host engine + C ABI (include/pasif_b200.h) of the B200-native RK4 forward / adjoint path.
//
// Replaces the reference's host driver `__GpuDriver` (reference src/__GpuDriver.cu) wholesale:
//   * hand-off analysis into component classes (pattern.hpp) instead of E-fold replication of the global system
//     (src/__GpuDriver.cu:896-938, 1196-1214),
//   * one persistent register-resident kernel per class for the whole forward sweep instead of a per-step
//     CUDA-graph of ~17 micro-kernels (src/__GpuDriver.cu:642-713),
//   * adjoint: checkpointed forward sweep, then per chunk one recompute launch + one backward launch
//     (src/__GpuDriver.cu:502-587, 715-783), with the reference's schedule reproduced literally in
//     compat=reference and an exact schedule in compat=clean,
//   * kernels are specialised on the sparsity pattern at compile time: prebuilt (patterns_aot.inc) for the
//     reference's example systems, NVRTC for any other pattern.  There is no interpreted or CPU fallback.
#include <cuda.h>
#include <cuda_runtime.h>
#include <dlfcn.h>
#include <sys/stat.h>
#include <unistd.h>
#include <time.h>

#include <algorithm>
#include <chrono>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <memory>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include "../../include/pasif_b200.h"
#include "pattern.hpp"
#include "lane_split.hpp"
#include "rk4_kernels.cuh"
#include "large_kernels.cuh"
#include "stage_kernels.cuh"
#include "stage_tcgen05.cuh"
#include "stage_stream.cuh"

using namespace pasif_host;

#define CUDA_CHECK(x)                                                                                          \
  do {                                                                                                         \
    cudaError_t _e = (x);                                                                                      \
    if (_e != cudaSuccess)                                                                                     \
      throw std::runtime_error(std::string("CUDA error ") + cudaGetErrorString(_e) + " at " + __FILE__ + ":" + \
                               std::to_string(__LINE__) + " (" #x ")");                                        \
  } while (0)

// ====================================================================================== prebuilt (AOT) kernels
namespace {
struct AotEntry {
  uint64_t hash;
  bool joint;
  const void* f64[3];  // forward | gradient-forward (dense save) | backward
  const void* f32[3];
  const void* f32x2[3];  // fp32, two trajectories per thread (packed FFMA2; rk4_kernels.cuh)
};
std::vector<AotEntry>& aot_table() {
  static std::vector<AotEntry> t;
  return t;
}
struct AotReg {
  explicit AotReg(const AotEntry& e) { aot_table().push_back(e); }
};
}  // namespace

#define PASIF_AOT_FWD(TR, HASH)                                                                                 \
  namespace {                                                                                                   \
  __global__ void __launch_bounds__(PASIF_BLOCK) pasif_fwd_f64_##TR(const pasif::FwdArgs<double> a) {           \
    pasif::forward_body<double, TR, pasif::NoDense>(a);                                                         \
  }                                                                                                             \
  __global__ void __launch_bounds__(PASIF_BLOCK) pasif_fwd_f32_##TR(const pasif::FwdArgs<float> a) {            \
    pasif::forward_body<float, TR, pasif::NoDense>(a);                                                          \
  }                                                                                                             \
  __global__ void __launch_bounds__(PASIF_BLOCK) pasif_fwd_f32x2_##TR(const pasif::FwdArgs<float> a) {          \
    pasif::forward_body<float, TR, pasif::NoDense, 2>(a);                                                       \
  }                                                                                                             \
  AotReg reg_##TR(AotEntry{HASH, false, {(const void*)pasif_fwd_f64_##TR, nullptr, nullptr},                    \
                           {(const void*)pasif_fwd_f32_##TR, nullptr, nullptr},                                 \
                           {(const void*)pasif_fwd_f32x2_##TR, nullptr, nullptr}});                             \
  }

#define PASIF_AOT_JOINT(TR, HASH)                                                                               \
  namespace {                                                                                                   \
  __global__ void __launch_bounds__(PASIF_BLOCK) pasif_gfwd_f64_##TR(const pasif::FwdArgs<double> a) {          \
    pasif::forward_body<double, TR::Fwd, TR::Adj>(a);                                                           \
  }                                                                                                             \
  __global__ void __launch_bounds__(PASIF_BLOCK) pasif_gfwd_f32_##TR(const pasif::FwdArgs<float> a) {           \
    pasif::forward_body<float, TR::Fwd, TR::Adj>(a);                                                            \
  }                                                                                                             \
  __global__ void __launch_bounds__(PASIF_BLOCK, 4) pasif_bwd_f64_##TR(const pasif::BwdArgs<double> a) {        \
    pasif::backward_body<double, TR::Adj>(a);                                                                   \
  }                                                                                                             \
  __global__ void __launch_bounds__(PASIF_BLOCK) pasif_bwd_f32_##TR(const pasif::BwdArgs<float> a) {            \
    pasif::backward_body<float, TR::Adj>(a);                                                                    \
  }                                                                                                             \
  __global__ void __launch_bounds__(PASIF_BLOCK) pasif_gfwd_f32x2_##TR(const pasif::FwdArgs<float> a) {         \
    pasif::forward_body<float, TR::Fwd, TR::Adj, 2>(a);                                                         \
  }                                                                                                             \
  __global__ void __launch_bounds__(PASIF_BLOCK, 4) pasif_bwd_f32x2_##TR(const pasif::BwdArgs<float> a) {       \
    pasif::backward_body<float, TR::Adj, 2>(a);                                                                 \
  }                                                                                                             \
  AotReg reg_##TR(AotEntry{HASH, true, {nullptr, (const void*)pasif_gfwd_f64_##TR, (const void*)pasif_bwd_f64_##TR}, \
                           {nullptr, (const void*)pasif_gfwd_f32_##TR, (const void*)pasif_bwd_f32_##TR},        \
                           {nullptr, (const void*)pasif_gfwd_f32x2_##TR, (const void*)pasif_bwd_f32x2_##TR}});  \
  }

#if __has_include("patterns_aot.inc")
#include "patterns_aot.inc"
#endif

// ====================================================================================== generic helper kernels
namespace {

// raw [E][L] double  ->  gated compute-type samples [rows][E]: row t holds exc[e][t] for 0 < t < L, else 0.
// The gate reproduces applyForces' `0 < step < L` (reference src/kernels.cu:82-87): a gated-out sample adds F*0.
// R: type of the staged samples -- double as handed over, or float when the host staged them in single precision (fp32
// engines: bit-identical, the cast happens on the host instead of here; fp64 engines: option exc_storage=f32)
template <typename T, typename R>
__global__ void prep_exc_full_kernel(const R* __restrict__ raw, T* __restrict__ out, int E, int L, int rows) {
  __shared__ R tile[32][33];
  const int t0 = blockIdx.x * 32, e0 = blockIdx.y * 32;
  for (int j = threadIdx.y; j < 32; j += blockDim.y) {
    const int e = e0 + j, t = t0 + threadIdx.x;
    tile[j][threadIdx.x] = (e < E && t > 0 && t < L) ? raw[(long long)e * L + t] : R(0);
  }
  __syncthreads();
  for (int j = threadIdx.y; j < 32; j += blockDim.y) {
    const int t = t0 + j, e = e0 + threadIdx.x;
    if (t < rows && e < E) out[(long long)t * E + e] = (T)tile[threadIdx.x][j];
  }
}

// interpolated half-step force stream f[u][e], u = 2*step + half (reference interpolateForces_fwd,
// src/kernels.cu:96-126): j = u & (W-1), o = u >> log2(W), 0 < o < L ? IM[W+j]*exc[o] + IM[j]*exc[o+1] : 0,
// exc[o+1] taken from the reference's flat [e][t] buffer (next excitation's first sample at o = L-1; 0 past the end).
template <typename T, typename R>
__global__ void prep_exc_half_kernel(const R* __restrict__ raw, const T* __restrict__ table, T* __restrict__ out,
                                     int E, int L, long long U, int W, int log2W) {
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= U * E) return;
  const long long u = idx / E;
  const int e = (int)(idx - u * E);
  const unsigned uu = (unsigned)u;
  const unsigned j = uu & (unsigned)(W - 1);
  const unsigned o = uu >> log2W;
  T v = T(0);
  if (o > 0 && o < (unsigned)L) {
    const long long pos = (long long)e * L + o;
    const T x0 = (T)raw[pos];
    const T x1 = (pos + 1 < (long long)E * L) ? (T)raw[pos + 1] : T(0);
    v = table[W + j] * x0 + table[j] * x1;
  }
  out[idx] = v;
}

// 2 x W table of linear interpolation between consecutive samples, generated on the device (the table a caller of the
// reference would build on the host and pass to setInterpolationMatrix; interpolateForces_fwd reads IM[W + j] as the weight
// of exc[o] and IM[j] as the weight of exc[o + 1], src/kernels.cu:114-124): IM[j] = j / W, IM[W + j] = 1 - j / W, evaluated
// in double and rounded to the compute type exactly like a host-built float64 table handed through the boundary.
template <typename T>
__global__ void gen_linear_table_kernel(T* __restrict__ table, int W) {
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= W) return;
  const double w = (double)j / (double)W;
  table[j] = (T)w;
  table[W + j] = (T)(1.0 - w);
}

// G = lanes per trajectory (1 for ordinary classes), W = trajectories per thread of the class's kernels (2: fp32 pair
// kernels).  Slot p of a trajectory-minor array belongs to thread p / W: lane (p / W) % G of trajectory
// ((p / W) / G) * W + p % W = inst * E + e; tables are indexed by the virtual instance inst * G + lane (rk4_kernels.cuh).
// For G == 1 this is trajectory p whatever W is; for W == 1 the r01 mapping lane p % G, trajectory p / G.
__device__ __forceinline__ void slot_to_lane_traj(int p, int G, int W, int& lane, int& traj) {
  const int t = p / W;
  lane = t % G;
  traj = (t / G) * W + p % W;
}
template <typename T>
__global__ void init_state_kernel(const T* __restrict__ ic, T* __restrict__ st, int N, int n_inst, int E, int n_traj, int G, int W) {
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (long long)N * n_traj) return;
  const int i = (int)(idx / n_traj);
  const int p = (int)(idx - (long long)i * n_traj);
  int lane, traj;
  slot_to_lane_traj(p, G, W, lane, traj);
  st[idx] = ic[(long long)i * n_inst + (traj / E) * G + lane];
}

template <typename T>
__global__ void scatter_user_kernel(const T* __restrict__ st, const int* __restrict__ varmap, T* __restrict__ out,
                                    int N, int n_inst, int E, int n_traj, int n_glob, int G, int W) {
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (long long)N * n_traj) return;
  const int i = (int)(idx / n_traj);
  const int p = (int)(idx - (long long)i * n_traj);
  int lane, traj;
  slot_to_lane_traj(p, G, W, lane, traj);
  const int e = traj % E, inst = (traj / E) * G + lane;
  const int g = varmap[(long long)i * n_inst + inst];
  if (g >= 0) out[(long long)e * n_glob + g] = st[idx];
}

// user layout [e][n_glob] -> class layout [N][n_traj] (the inverse of scatter_user_kernel); padding slots take the class's
// initial condition.  Used for the call-order state (a sweep that starts from the state a previous call left behind).
template <typename T>
__global__ void gather_user_kernel(const T* __restrict__ user, const int* __restrict__ varmap, const T* __restrict__ ic,
                                   T* __restrict__ st, int N, int n_inst, int E, int n_traj, int n_glob, int G, int W) {
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (long long)N * n_traj) return;
  const int i = (int)(idx / n_traj);
  const int p = (int)(idx - (long long)i * n_traj);
  int lane, traj;
  slot_to_lane_traj(p, G, W, lane, traj);
  const int e = traj % E, inst = (traj / E) * G + lane;
  const int g = varmap[(long long)i * n_inst + inst];
  st[idx] = g >= 0 ? user[(long long)e * n_glob + g] : ic[(long long)i * n_inst + inst];
}

// out[j] = sum_e in[e][j]: each warp owns 32 consecutive columns while striding over e, the 32 row-partials of a
// column are then transposed through shared memory and folded with warp shuffles (fixed order -> deterministic).
template <typename T>
__global__ void reduce_excitations_kernel(const T* __restrict__ in, T* __restrict__ out, int E, int n) {
  __shared__ T part[32][33];
  const int j = blockIdx.x * 32 + threadIdx.x;
  T acc = T(0);
  if (j < n)
    for (int e = threadIdx.y; e < E; e += 32) acc += in[(long long)e * n + j];
  part[threadIdx.y][threadIdx.x] = acc;
  __syncthreads();
  T v = part[threadIdx.x][threadIdx.y];          // warp `y` now holds the 32 partials of column blockIdx.x*32 + y
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) v += __shfl_down_sync(0xffffffffu, v, off);
  const int jo = blockIdx.x * 32 + threadIdx.y;
  if (threadIdx.x == 0 && jo < n) out[jo] = v;
}

// register-resident FMA-pipe peak: 16 independent chains per thread, 8 rounds per loop iteration = 128 FMAs against
// 3 loop instructions (cuobjdump: the loop body is 128 FFMA / DFMA + IADD3 + ISETP + BRA, 97.7 % FMA; the r01 version
// carried 8 FMAs per iteration and under-read the pipe by 9 %).  No memory traffic in the loop.
template <typename T>
__global__ void fma_peak_kernel(T* out, int iters, T a, T b) {
  T x[16];
#pragma unroll
  for (int k = 0; k < 16; ++k) x[k] = T(threadIdx.x + k);
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int r = 0; r < 8; ++r)
#pragma unroll
      for (int k = 0; k < 16; ++k) x[k] = fma(x[k], a, b);
  }
  T s = T(0);
#pragma unroll
  for (int k = 0; k < 16; ++k) s += x[k];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
// the same with packed fp32 pairs (FFMA2): two FMAs per instruction slot -- the pipe's peak does not change, the issue
// rate needed to reach it halves
__global__ void fma2_peak_kernel(float2* out, int iters, float a, float b) {
  float2 x[16];
  const float2 a2 = make_float2(a, a), b2 = make_float2(b, b);
#pragma unroll
  for (int k = 0; k < 16; ++k) x[k] = make_float2(float(threadIdx.x + k), float(k));
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int r = 0; r < 8; ++r)
#pragma unroll
      for (int k = 0; k < 16; ++k) x[k] = __ffma2_rn(x[k], a2, b2);
  }
  float2 s = make_float2(0.f, 0.f);
#pragma unroll
  for (int k = 0; k < 16; ++k) { s.x += x[k].x; s.y += x[k].y; }
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// legacy tensor path peaks (mma.sync), 4 independent accumulator chains per warp
__global__ void dmma_peak_kernel(double* out, int iters) {
  double c[4][2] = {{0, 0}, {0, 0}, {0, 0}, {0, 0}};
  const double a = 1.0 + threadIdx.x * 1e-9, b = 1.0 - threadIdx.x * 1e-9;
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int k = 0; k < 4; ++k)
      asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                   : "+d"(c[k][0]), "+d"(c[k][1]) : "d"(a), "d"(b));
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = c[0][0] + c[1][1] + c[2][0] + c[3][1];
}
__global__ void tf32_mma_peak_kernel(float* out, int iters) {
  float c[4][4] = {};
  const unsigned a = __float_as_uint(1.0f), b = __float_as_uint(0.5f);
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int k = 0; k < 4; ++k)
      asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                   : "+f"(c[k][0]), "+f"(c[k][1]), "+f"(c[k][2]), "+f"(c[k][3]) : "r"(a), "r"(a), "r"(a), "r"(a), "r"(b), "r"(b));
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = c[0][0] + c[1][1] + c[2][2] + c[3][3];
}

}  // namespace

// ====================================================================================== NVRTC / driver entry points
namespace {

// process-global lazy state (NVRTC / driver entry points, the kernel header text) is initialised under one mutex: the
// pybind shim releases the GIL while systems are allocated, so two engines may get here from two threads
std::mutex g_init_mu;

typedef struct _nvrtcProgram* nvrtcProgram;
struct Nvrtc {
  void* lib = nullptr;
  int (*CreateProgram)(nvrtcProgram*, const char*, const char*, int, const char* const*, const char* const*) = nullptr;
  int (*CompileProgram)(nvrtcProgram, int, const char* const*) = nullptr;
  int (*GetCUBINSize)(nvrtcProgram, size_t*) = nullptr;
  int (*GetCUBIN)(nvrtcProgram, char*) = nullptr;
  int (*GetProgramLogSize)(nvrtcProgram, size_t*) = nullptr;
  int (*GetProgramLog)(nvrtcProgram, char*) = nullptr;
  int (*DestroyProgram)(nvrtcProgram*) = nullptr;
  int (*Version)(int*, int*) = nullptr;
  void load() {
    std::lock_guard<std::mutex> lk(g_init_mu);
    if (lib && Version) return;
    const char* names[] = {"libnvrtc.so.12", "libnvrtc.so", "/usr/local/cuda/lib64/libnvrtc.so.12",
                           "/usr/local/cuda/lib64/libnvrtc.so"};
    for (const char* n : names)
      if ((lib = dlopen(n, RTLD_NOW | RTLD_LOCAL))) break;
    if (!lib) throw std::runtime_error("NVRTC (libnvrtc.so) could not be loaded: no prebuilt kernel matches this "
                                       "system's sparsity pattern and run-time specialisation is unavailable");
    auto sym = [&](const char* s) {
      void* p = dlsym(lib, s);
      if (!p) throw std::runtime_error(std::string("NVRTC symbol missing: ") + s);
      return p;
    };
    CreateProgram = (decltype(CreateProgram))sym("nvrtcCreateProgram");
    CompileProgram = (decltype(CompileProgram))sym("nvrtcCompileProgram");
    GetCUBINSize = (decltype(GetCUBINSize))sym("nvrtcGetCUBINSize");
    GetCUBIN = (decltype(GetCUBIN))sym("nvrtcGetCUBIN");
    GetProgramLogSize = (decltype(GetProgramLogSize))sym("nvrtcGetProgramLogSize");
    GetProgramLog = (decltype(GetProgramLog))sym("nvrtcGetProgramLog");
    DestroyProgram = (decltype(DestroyProgram))sym("nvrtcDestroyProgram");
    Version = (decltype(Version))sym("nvrtcVersion");
  }
};
Nvrtc g_nvrtc;
// process-global lazy state (NVRTC / driver entry points, the kernel header text) is initialised under one mutex: the
// pybind shim releases the GIL while systems are allocated, so two engines may get here from two threads

struct DriverApi {
  bool ok = false;
  CUresult (*ModuleLoadData)(CUmodule*, const void*) = nullptr;
  CUresult (*ModuleGetFunction)(CUfunction*, CUmodule, const char*) = nullptr;
  CUresult (*ModuleUnload)(CUmodule) = nullptr;
  CUresult (*LaunchKernel)(CUfunction, unsigned, unsigned, unsigned, unsigned, unsigned, unsigned, unsigned, CUstream,
                           void**, void**) = nullptr;
  void load() {
    std::lock_guard<std::mutex> lk(g_init_mu);
    if (ok) return;
    auto get = [&](const char* name, void** fn) {
      cudaDriverEntryPointQueryResult q;
      CUDA_CHECK(cudaGetDriverEntryPoint(name, fn, cudaEnableDefault, &q));
      if (q != cudaDriverEntryPointSuccess || !*fn) throw std::runtime_error(std::string("driver entry point missing: ") + name);
    };
    get("cuModuleLoadData", (void**)&ModuleLoadData);
    get("cuModuleGetFunction", (void**)&ModuleGetFunction);
    get("cuModuleUnload", (void**)&ModuleUnload);
    get("cuLaunchKernel", (void**)&LaunchKernel);
    ok = true;
  }
};
DriverApi g_drv;

std::string read_file(const std::string& path) {
  std::ifstream f(path, std::ios::binary);
  if (!f) return std::string();
  return std::string((std::istreambuf_iterator<char>(f)), std::istreambuf_iterator<char>());
}

// text of rk4_kernels.cuh for run-time specialisation: compiled into the library (rk4_kernels_text.inc, written by
// pasif_b200/build.py), so a deployed .so needs no source file beside it; PASIF_B200_KERNEL_HEADER overrides it (kernel
// development without rebuilding the library)
const char* const kEmbeddedKernelHeader =
#include "rk4_kernels_text.inc"
    ;

std::string kernel_header_text() {
  static std::string cached;
  std::lock_guard<std::mutex> lk(g_init_mu);
  if (!cached.empty()) return cached;
  if (const char* env = getenv("PASIF_B200_KERNEL_HEADER")) {
    cached = read_file(env);
    if (cached.empty()) throw std::runtime_error(std::string("PASIF_B200_KERNEL_HEADER: cannot read ") + env);
    return cached;
  }
  cached = kEmbeddedKernelHeader;
  return cached;
}

struct KernelHandle {
  const void* aot = nullptr;
  CUfunction jit = nullptr;
  bool valid() const { return aot || jit; }
};

}  // namespace

// ====================================================================================== engine
namespace {

struct DevBuf;
// buffer that may be given up when another allocation of the same engine runs out of device memory: the chunk
// trajectories the exact schedule keeps in otherwise free HBM (d_tail).  Set by pasif_engine::use_device().
thread_local DevBuf* tl_oom_victim = nullptr;

struct DevBuf {
  void* p = nullptr;
  size_t cap = 0;
  void reserve(size_t bytes) {
    if (bytes <= cap) return;
    if (p) cudaFree(p);
    p = nullptr; cap = 0;
    cudaError_t st = cudaMalloc(&p, bytes);
    if (st == cudaErrorMemoryAllocation && tl_oom_victim && tl_oom_victim != this && tl_oom_victim->p) {
      (void)cudaGetLastError();
      tl_oom_victim->release();                  // kept chunk trajectories are a cache: the next gradient call re-sizes them
      st = cudaMalloc(&p, bytes);
    }
    if (st != cudaSuccess) { p = nullptr; CUDA_CHECK(st); }
    cap = bytes;
  }
  void release() {
    if (p) cudaFree(p);
    p = nullptr; cap = 0;
  }
};

// page-locked host staging area: user buffers are pageable (NumPy arrays, std::vector); copying through two pinned
// slots keeps the DMA engine busy while the host thread fills / drains the other slot
struct PinnedBuf {
  void* p = nullptr;
  size_t cap = 0;
  void reserve(size_t bytes) {
    if (bytes <= cap) return;
    if (p) cudaFreeHost(p);
    p = nullptr; cap = 0;
    CUDA_CHECK(cudaHostAlloc(&p, bytes, cudaHostAllocDefault));
    cap = bytes;
  }
  void release() {
    if (p) cudaFreeHost(p);
    p = nullptr; cap = 0;
  }
};

struct LargeClass;
struct LargeClassDeleter { void operator()(LargeClass* p) const; };

struct DevClass {
  ClassData h;
  std::unique_ptr<LargeClass, LargeClassDeleter> large;   // set for components beyond the register-resident range
  DevBuf coef, ic, varmap, coef_adj, ic_adj, varmap_adj;
  KernelHandle k_fwd, k_gfwd, k_bwd;
  KernelHandle k_fwd2, k_gfwd2, k_bwd2;         // fp32 one-thread classes: two trajectories per thread (packed FFMA2)
  DevBuf lam_carry;                             // call_order=reference: adjoint state the last gradient call left behind [NA][n_traj]
  bool lam_dirty = false;
  CUmodule module = nullptr;
  long long n_traj = 0;
};

}  // namespace

namespace {
struct StageBufs;   // work arrays of the dense-operator path (stage_host.inc)
}

struct pasif_engine {
  int dtype = PASIF_F64, device = 0;
  StageBufs* stage_bufs[2] = {nullptr, nullptr};
  bool host_only = false;
  uint32_t E = 0, L = 0, sample_rate = 1, numsteps = 0;
  bool have_exc = false;
  SystemData fwd, bwd;
  std::vector<double> interp_table, modulation;
  uint32_t interp_window = 0;
  bool interp_linear = false;                      // the 2 x W table is the linear one, generated on the device
  bool exc_storage_f32 = false;                    // fp64 engines: stage / keep the excitation samples in single precision
  bool raw_f32 = false;                            // what d_raw currently holds
  std::vector<DevClass> fwd_classes, joint_classes;
  bool fwd_ready = false, joint_ready = false;
  bool exc_dirty = true;
  DevBuf d_tail;                                  // clean schedule: chunk trajectories kept from the forward sweep (sized to HBM)
  double tail_budget_gb = -1;                     // < 0: automatic (a fraction of the free device memory); 0: off
  long long last_stored_chunks = 0, last_chunks = 0;
  DevBuf d_raw, d_exc, d_user, d_traj, d_ckpt, d_dense, d_dense2, d_lam, d_red;   // d_raw: excitations as handed over, [E][L] doubles
  PinnedBuf h_stage;                               // 2 slots of kStageSlot bytes
  cudaEvent_t ev_stage[2] = {nullptr, nullptr};
  // Call-order state of the reference driver (option call_order = reference | independent).  The reference
  // resets its forward state d_fwd_Q to the initial conditions only AFTER getAmplitudes / getTrajectory and when the system
  // is allocated (src/__GpuDriver.cu:447, 482, 1183-1186); _getGradient starts from whatever it holds (:522-523) and leaves
  // the end of its last recomputed chunk behind (:565), and the adjoint initial-condition buffer is overwritten with the
  // running adjoint state (:743).  Reproduced for the register / lane-block classes in compat=reference: d_fwd_cur holds the
  // forward state a previous gradient call left behind (user layout), DevClass::lam_carry the adjoint state.
  bool call_order_ref = true;
  bool fwd_dirty = false;
  DevBuf d_fwd_cur, d_q0, d_qend;                  // [E][N_glob] | class-layout start state | class-layout end state
  int d_user_kind = 0;                             // what d_user holds: 0 nothing | 1 amplitudes [E][N] | 2 adjoint state [E][N_adj]
  double last_h2d_ms = 0, last_d2h_ms = 0;
  size_t last_h2d_bytes = 0;                       // bytes the last excitation upload moved over the host link         // host wall time of the last staged upload / download
  long long exc_rows = 0;
  bool exc_half = false;
  cudaStream_t stream = nullptr, stream2 = nullptr;   // stream2: adjoint sweep of chunk k overlaps the recompute of chunk k-1
  cudaEvent_t ev0 = nullptr, ev1 = nullptr;
  cudaEvent_t ev_rec[2] = {nullptr, nullptr}, ev_bwd[2] = {nullptr, nullptr};
  // dense-operator path: the batch is cut into `stage_nparts` parts, each with its own stream, so that the tensor-bound
  // GEMM of one part overlaps the HBM-bound update of another (part 0 runs on `stream`)
  int stage_nparts = 2;
  int update_u = 4;                              // trajectories per thread of the dense-operator path's update kernel
  bool update_stream = true;                     // dense-operator path: persistent bulk-copy-pipelined update kernel (stage_stream.cuh)
  int stream_tt = 0;                             // ... trajectories per tile (0: 4 in fp32, 2 in fp64)
  int stream_ctas = 2;                           // ... resident CTAs per SM the grid is sized for
  int stream_smem_kb = 0;                        // ... shared memory of one CTA (0: the SM's share); with tc_stages: room for a GEMM CTA beside it
  int stream_grid = 0;                           // ... CTAs per SM in the grid (0: stream_ctas)
  int stream_grid_ctas = 0;                      // ... or the absolute number of CTAs (0: per SM as above)
  bool stream_dynamic = false;                   // ... tiles drawn from a counter instead of a fixed share per CTA
  // dense-operator path: the launch sequence of a gradient as an instantiated CUDA graph (stage_host.inc)
  struct StageGraphT {
    cudaGraphExec_t exec = nullptr;
    std::vector<uint64_t> key, warm_key;
    long long d_launches = 0, d_gemm = 0, d_kind[5] = {0, 0, 0, 0, 0};
    long long warm_launches = 0;                  // launches of the directly queued call with warm_key
  } stage_graph;
  bool stage_graph_on = true;
  long long stage_graph_replays = 0;
  uint64_t cfg_gen = 0;                          // bumped by every option / system / excitation change: invalidates the graph
  long long stage_enqueue_us = 0;                // host time the last dense-operator gradient spent queueing launches
  DevBuf stage_sched;                            // ... the counters: (next tile, finished CTAs) per part
  int dg_bm = 32;                                // fp64 stage GEMM (DMMA): rows per CTA tile (32: four CTAs of 4 warps per SM; 64: two of 8)
  int tc_stages = 0;                             // fused-split GEMM: pipeline stages (0: as many as fit an SM)
  bool stage_dephase = false;                    // parts start one GEMM apart, so GEMMs run beside the neighbour's updates
  std::vector<cudaStream_t> part_streams;
  std::vector<cudaEvent_t> part_events;
  bool overlap = true;
  std::string err;
  bool compat_clean = false, jit_on = true, gemm_on = true;
  int large_threshold = 48;                      // components with more coupled variables use the lane-block / tile kernels
  int lanes_min = 8;                             // narrowest thread group the split tries (2 / 4: small systems of a small batch, see DESIGN "Small batches")
  int lanes_max = 1024;                          // widest thread group of the lane-block split (0: off)
  bool unit_on = true;                           // register kernels: K terms with coefficient exactly 1 become additions
  bool pair_on = true;                           // fp32 one-thread classes: two trajectories per thread (packed FFMA2)
  bool lane_pair_on = true;                      // lane-block classes of an fp32 engine: pair kernels, split sized for them
  int pair_regs = 230;                           // register estimate up to which the pair kernels are used
  long long pair_launches = 0;
  bool aot_on = true;                            // use the prebuilt kernels when a pattern matches (off: always specialise at run time)
  int jit_minb_fwd = 0, jit_minb_bwd = 0;        // one-thread kernels specialised at run time: minimum CTAs per SM (0: none)
  int lane_minb_fwd = 0, lane_minb_bwd = 0;      // lane-block kernels: minimum CTAs per SM handed to __launch_bounds__ (0: none)
  int one_thread_regs = 330;                     // register estimate above which a component is offered to the lane-block split
  double lane_min_eff = 0.30;                    // smallest share of real terms in the common lane pattern that is accepted
  int lane_regs = 240;                           // register estimate a lane pattern may reach before a wider group is tried
  long long lane_classes = 0, lane_selfcheck_failures = 0;
  LaneSplitReport lane_report;                   // of the last class that was split
  std::string lane_reject;                       // why the last large class was not split
  long long large_classes = 0, large_gemm_classes = 0, large_stage_classes = 0, gemm_launches = 0;
  bool tcgen05_on = true;                        // fp32 stage GEMM on tcgen05/TMEM/TMA (else the mma.sync 3xTF32 kernel)
  bool tc_pair = true;                           // fused-split GEMM on CTA pairs (tcgen05.mma.cta_group::2: every CTA holds half of the K' tile)
  int tc_pair_nt = 2;                            // ... row tiles of 256 per pair (2 when the band is a multiple of 512 rows)
  bool tc_multicast = false;                     // fused-split GEMM on clusters of two CTAs with the K' tile multicast (experimental)
  bool stage_direct = true;                      // dense-operator path: the state array is the stage vector of stage 1 (stage 4 writes one array)
  int tc_bk = 16;                                // fused-split GEMM: floats per k-block (16: 64-byte swizzle, twice the pipeline stages; 32: r01)
  int tc_bn = 256;                               // fused-split GEMM: 256-row tiles when the band of dense rows is a multiple of 256 (else 128 / 192)
  bool tc_fused_split = true;                    // tcgen05 GEMM splits the fp32 stage vector into tf32 hi / lo in shared memory
  bool gemm_band = true;                         // dense-operator path: only the band of dense rows goes through the GEMM
  bool gemm_stage = true;                        // dense shared K: per-stage batched GEMM with HBM-resident state (else in-tile GEMM)
  int block = PASIF_BLOCK, n_sm = 148;
  double last_ms = 0;
  long long last_launches = 0, jit_compiles = 0, aot_hits = 0, jit_cache_hits = 0;
  // option profile=on: CUDA-event pair around every hot-kernel launch, summed per kind (0 fwd, 1 gradient-fwd, 2 bwd)
  bool profile = false;
  std::vector<cudaEvent_t> prof_events;
  std::vector<int> prof_kinds;
  size_t prof_used = 0;
  // profile=on: device time per kernel kind (0 forward, 1 gradient sweep / recompute, 2 adjoint; dense-operator path: those
  // are its GEMMs, and 3 / 4 its update kernels of the forward system / of the adjoint)
  double kind_ms[5] = {0, 0, 0, 0, 0};
  long long kind_launches[5] = {0, 0, 0, 0, 0};

  size_t tsize() const { return dtype == PASIF_F32 ? 4 : 8; }
  void use_device() {
    if (host_only) throw std::runtime_error("this engine was created host-only (device = -1): no CUDA device to solve on");
    CUDA_CHECK(cudaSetDevice(device));
    tl_oom_victim = &d_tail;
  }
};

static thread_local std::string g_create_error;   // pasif_create / the handle-less debug calls report per calling thread

namespace {

template <typename T>
void upload(DevBuf& b, const std::vector<double>& v) {
  std::vector<T> t(v.begin(), v.end());
  b.reserve(std::max<size_t>(t.size(), 1) * sizeof(T));
  if (!t.empty()) CUDA_CHECK(cudaMemcpy(b.p, t.data(), t.size() * sizeof(T), cudaMemcpyHostToDevice));
}
void upload_int(DevBuf& b, const std::vector<int>& v) {
  b.reserve(std::max<size_t>(v.size(), 1) * sizeof(int));
  if (!v.empty()) CUDA_CHECK(cudaMemcpy(b.p, v.data(), v.size() * sizeof(int), cudaMemcpyHostToDevice));
}

// ---- kernel lookup: prebuilt table first, NVRTC otherwise
std::string jit_source(const ClassData& c, int dtype, int minb_fwd = 0, int minb_bwd = 0) {
  // lane groups wider than the default CTA are compiled for a CTA of their own width
  std::string s = c.lanes > PASIF_BLOCK ? "#define PASIF_BLOCK " + std::to_string(c.lanes) + "\n" : std::string();
  s += kernel_header_text();
  s += "\n";
  s += emit_class_traits(c);
  const char* T = dtype == PASIF_F32 ? "float" : "double";
  // minb > 0: __launch_bounds__(PASIF_BLOCK, minb) -- caps the registers so that `minb` CTAs fit an SM (lane-block classes)
  auto lb = [](int minb) { return minb > 0 ? "__launch_bounds__(PASIF_BLOCK, " + std::to_string(minb) + ")" : std::string("__launch_bounds__(PASIF_BLOCK)"); };
  const bool pair = dtype == PASIF_F32 && (c.lanes == 1 || c.lane_pair);
  if (pair) {
    if (!c.joint)
      s += "extern \"C\" __global__ void " + lb(minb_fwd) + " pasif_k_fwd2(const pasif::FwdArgs<float> a) { pasif::forward_body<float, " +
           c.name + ", pasif::NoDense, 2>(a); }\n";
    else {
      s += "extern \"C\" __global__ void " + lb(minb_fwd) + " pasif_k_gfwd2(const pasif::FwdArgs<float> a) { pasif::forward_body<float, " +
           c.name + "::Fwd, " + c.name + "::Adj, 2>(a); }\n";
      // the adjoint pair kernel of the 12-state example Jacobian needs 150 registers uncapped = 3 CTAs per SM; capped at 128
      // (4 CTAs) it runs 4 % faster (profiles/r02_probe_pair_variants.json).  Larger patterns keep the uncapped build.
      const int est = c.adj.nc() + 8 * c.adj.N + 8 * (int)c.adj.nv.size() + 24;
      s += "extern \"C\" __global__ void " + lb(minb_bwd > 0 ? minb_bwd : ((c.lanes == 1 && est <= 180) ? 4 : 0)) + " pasif_k_bwd2(const pasif::BwdArgs<float> a) { pasif::backward_body<float, " +
           c.name + "::Adj, 2>(a); }\n";
    }
  }
  if (!c.joint) {
    s += std::string("extern \"C\" __global__ void ") + lb(minb_fwd) + " pasif_k_fwd(const pasif::FwdArgs<" + T +
         "> a) { pasif::forward_body<" + T + ", " + c.name + ", pasif::NoDense>(a); }\n";
  } else {
    s += std::string("extern \"C\" __global__ void ") + lb(minb_fwd) + " pasif_k_gfwd(const pasif::FwdArgs<" + T +
         "> a) { pasif::forward_body<" + T + ", " + c.name + "::Fwd, " + c.name + "::Adj>(a); }\n";
    // fp64 adjoint kernel of the 12-state example Jacobian: 162 registers uncapped = 3 CTAs per SM; capped at 128 (4 CTAs, no
    // spills) it runs 5 % faster.  Same rule as for the fp32 pair kernel above; larger patterns keep the uncapped build.
    const int est64 = 2 * (c.adj.nc() + 4 * c.adj.N + 4 * (int)c.adj.nv.size()) + 24;
    const int cap_bwd = minb_bwd > 0 ? minb_bwd : ((dtype == PASIF_F64 && c.lanes == 1 && est64 <= 200) ? 4 : 0);
    s += std::string("extern \"C\" __global__ void ") + lb(cap_bwd) + " pasif_k_bwd(const pasif::BwdArgs<" + T +
         "> a) { pasif::backward_body<" + T + ", " + c.name + "::Adj>(a); }\n";
  }
  return s;
}

// on-disk cache of the run-time specialised cubins.  Cubins are loaded as GPU code, so the directory must belong to this
// user and be closed to others (0700); with neither PASIF_B200_CACHE nor HOME there is no cache (never a shared /tmp path).
std::string cache_dir() {
  std::string dir;
  if (const char* e = getenv("PASIF_B200_CACHE")) dir = e;
  else if (const char* home = getenv("HOME")) {
    mkdir((std::string(home) + "/.cache").c_str(), 0755);
    dir = std::string(home) + "/.cache/pasif_b200";
  } else return std::string();
  mkdir(dir.c_str(), 0700);
  struct stat st;
  if (stat(dir.c_str(), &st) != 0 || !S_ISDIR(st.st_mode) || st.st_uid != geteuid() || (st.st_mode & 077) != 0) return std::string();
  return dir;
}

std::vector<char> jit_compile(pasif_engine& E, const std::string& src) {
  g_nvrtc.load();
  cudaDeviceProp prop;
  CUDA_CHECK(cudaGetDeviceProperties(&prop, E.device));
  std::string arch = "--gpu-architecture=sm_" + std::to_string(prop.major) + std::to_string(prop.minor) + (prop.major >= 9 ? "a" : "");
  int vmaj = 0, vmin = 0;
  g_nvrtc.Version(&vmaj, &vmin);
  const std::string key = detail::hex64(fnv1a(src + arch + std::to_string(vmaj * 100 + vmin)));
  const std::string dir = cache_dir(), path = dir + "/" + key + ".cubin";
  if (!dir.empty()) {
    std::string blob = read_file(path);
    if (!blob.empty()) { ++E.jit_cache_hits; return std::vector<char>(blob.begin(), blob.end()); }
  }
  nvrtcProgram prog = nullptr;
  if (g_nvrtc.CreateProgram(&prog, src.c_str(), "pasif_jit.cu", 0, nullptr, nullptr) != 0)
    throw std::runtime_error("nvrtcCreateProgram failed");
  const char* opts[] = {arch.c_str(), "-std=c++17", "-lineinfo"};
  int rc = g_nvrtc.CompileProgram(prog, 3, opts);
  if (rc != 0) {
    size_t n = 0;
    g_nvrtc.GetProgramLogSize(prog, &n);
    std::string log(n, '\0');
    if (n) g_nvrtc.GetProgramLog(prog, &log[0]);
    g_nvrtc.DestroyProgram(&prog);
    throw std::runtime_error("NVRTC compilation of the specialised kernel failed:\n" + log);
  }
  size_t n = 0;
  g_nvrtc.GetCUBINSize(prog, &n);
  std::vector<char> cubin(n);
  g_nvrtc.GetCUBIN(prog, cubin.data());
  g_nvrtc.DestroyProgram(&prog);
  ++E.jit_compiles;
  // best-effort disk cache (private directory only, see cache_dir)
  if (dir.empty()) return cubin;
  std::ofstream f(path + "." + std::to_string((long long)getpid()) + ".tmp", std::ios::binary);
  if (f) {
    f.write(cubin.data(), (std::streamsize)cubin.size());
    f.close();
    rename((path + "." + std::to_string((long long)getpid()) + ".tmp").c_str(), path.c_str());
  }
  return cubin;
}

void resolve_kernels(pasif_engine& E, DevClass& dc) {
  const ClassData& c = dc.h;
  for (const AotEntry& a : aot_table()) {
    if (!E.aot_on) break;
    if (a.hash == c.hash && a.joint == c.joint) {
      const void* const* f = E.dtype == PASIF_F32 ? a.f32 : a.f64;
      dc.k_fwd.aot = f[0]; dc.k_gfwd.aot = f[1]; dc.k_bwd.aot = f[2];
      if (E.dtype == PASIF_F32) { dc.k_fwd2.aot = a.f32x2[0]; dc.k_gfwd2.aot = a.f32x2[1]; dc.k_bwd2.aot = a.f32x2[2]; }
      ++E.aot_hits;
      return;
    }
  }
  if (!E.jit_on)
    throw std::runtime_error("no prebuilt kernel for pattern " + c.name + " and option jit=off");
  const int nmax = std::max(c.fwd.N, c.adj.N);
  if (nmax > 48 || c.fwd.nc() + c.adj.nc() > 2048)
    throw std::runtime_error("component " + c.name + " has " + std::to_string(nmax) + " coupled state variables / " +
                             std::to_string(c.fwd.nc() + c.adj.nc()) + " coefficients: beyond the one-thread register "
                             "kernel's range (<= 48 variables, <= 2048 coefficients) -- internal: such a component belongs to "
                             "the lane-block / dense-operator / tile paths");
  g_drv.load();
  std::vector<char> cubin = jit_compile(E, jit_source(c, E.dtype, c.lanes > 1 ? E.lane_minb_fwd : E.jit_minb_fwd, c.lanes > 1 ? E.lane_minb_bwd : E.jit_minb_bwd));
  if (g_drv.ModuleLoadData(&dc.module, cubin.data()) != CUDA_SUCCESS)
    throw std::runtime_error("cuModuleLoadData failed for the specialised kernel of " + c.name);
  auto getf = [&](const char* n, CUfunction* f) {
    if (g_drv.ModuleGetFunction(f, dc.module, n) != CUDA_SUCCESS)
      throw std::runtime_error(std::string("cuModuleGetFunction failed: ") + n);
  };
  if (!c.joint) getf("pasif_k_fwd", &dc.k_fwd.jit);
  else { getf("pasif_k_gfwd", &dc.k_gfwd.jit); getf("pasif_k_bwd", &dc.k_bwd.jit); }
  if (E.dtype == PASIF_F32 && (c.lanes == 1 || c.lane_pair)) {
    if (!c.joint) getf("pasif_k_fwd2", &dc.k_fwd2.jit);
    else { getf("pasif_k_gfwd2", &dc.k_gfwd2.jit); getf("pasif_k_bwd2", &dc.k_bwd2.jit); }
  }
}

cudaEvent_t prof_event(pasif_engine& E, int kind) {
  if (E.prof_used == E.prof_events.size()) {
    cudaEvent_t ev;
    CUDA_CHECK(cudaEventCreate(&ev));
    E.prof_events.push_back(ev);
  }
  if (E.prof_used % 2 == 0) E.prof_kinds.push_back(kind);
  return E.prof_events[E.prof_used++];
}

// threads per CTA: the configured maximum while the grid still gives every SM >= 2 CTAs, else narrower CTAs so that
// a small batch (config C2: 1024 trajectories) spreads over more SMs
unsigned pick_block(const pasif_engine& E, long long n_threads, unsigned min_block = 32) {
  unsigned block = (unsigned)E.block;
  while (block > min_block && block > 32 && (n_threads + block - 1) / block < 2ll * E.n_sm) block >>= 1;
  return block < 32 ? 32 : block;
}

template <typename Args>
void launch(pasif_engine& E, const KernelHandle& k, long long n_threads, Args& args, int kind, cudaStream_t st = nullptr,
            int lanes = 1) {
  if (!k.valid()) throw std::runtime_error("internal: kernel handle not resolved");
  if (!st) st = E.stream;
  // lane-block classes: a group of `lanes` threads must sit in one CTA; they are compiled for the full block
  const unsigned block = lanes > 1 ? (unsigned)std::max(PASIF_BLOCK, lanes) : pick_block(E, n_threads);
  const unsigned grid = (unsigned)((n_threads + block - 1) / block);
  void* params[] = {(void*)&args};
  if (E.profile) CUDA_CHECK(cudaEventRecord(prof_event(E, kind), st));
  if (k.aot) CUDA_CHECK(cudaLaunchKernel(k.aot, dim3(grid), dim3(block), params, 0, st));
  else if (g_drv.LaunchKernel(k.jit, grid, 1, 1, block, 1, 1, 0, (CUstream)st, params, nullptr) != CUDA_SUCCESS)
    throw std::runtime_error("cuLaunchKernel failed");
  if (E.profile) CUDA_CHECK(cudaEventRecord(prof_event(E, kind), st));
  ++E.last_launches;
  ++E.kind_launches[kind];
}

// fp32 one-thread classes with an even number of excitations run TWO trajectories per thread (packed FFMA2 kernels,
// rk4_kernels.cuh): the pair (e, e+1) belongs to one instance and every 8-byte access is aligned.  Bit-identical to the
// one-trajectory kernels; option pair=off keeps those.
bool use_pairs(const pasif_engine& E, const DevClass& c) {
  if (!E.pair_on || E.dtype != PASIF_F32 || c.large || E.E % 2 != 0) return false;
  if (!(c.h.joint ? c.k_gfwd2.valid() && c.k_bwd2.valid() : c.k_fwd2.valid())) return false;
  // lane-block classes: only when the split was sized for pairs (lane_split.hpp: state slots counted twice)
  if (c.h.lanes > 1) return c.h.lane_pair;
  // register estimate of the pair kernels (calibrated with ptxas -v: 8-state system 100, 12-state adjoint 152): coefficients
  // stay scalars, every state slot doubles; beyond the register file's comfortable range the one-trajectory kernels stay
  const long long est = std::max<long long>(c.h.fwd.nc() + 8ll * c.h.fwd.N,
                                            c.h.joint ? c.h.adj.nc() + 8ll * c.h.adj.N + 8ll * (long long)c.h.adj.nv.size() : 0) + 24;
  return est <= E.pair_regs;
}

void free_classes(std::vector<DevClass>& v) {
  for (auto& c : v) {
    c.coef.release(); c.ic.release(); c.varmap.release();
    c.coef_adj.release(); c.ic_adj.release(); c.varmap_adj.release();
    c.lam_carry.release();
    c.large.reset();
    if (c.module && g_drv.ok) g_drv.ModuleUnload(c.module);
  }
  v.clear();
}

template <typename T>
void build_large_class(pasif_engine& E, DevClass& dc);
template <typename T>
void large_solve_forward(pasif_engine& E, DevClass& c, bool trajectory, uint32_t save_steps, uint64_t n_slots);
template <typename T>
void large_solve_gradient(pasif_engine& E, DevClass& c, uint32_t save, T* d_user_out);
bool is_large(const pasif_engine& E, const ClassData& c) {
  return std::max(c.fwd.N, c.adj.N) > E.large_threshold || c.fwd.nc() + c.adj.nc() > 2048;
}

template <typename T>
void build_classes(pasif_engine& E, std::vector<ClassData>&& cds, std::vector<DevClass>& out) {
  ++E.cfg_gen;
  free_classes(out);
  E.d_tail.release();                              // sized for the previous systems
  out.resize(cds.size());
  for (size_t i = 0; i < cds.size(); ++i) {
    DevClass& dc = out[i];
    dc.h = std::move(cds[i]);
    // offered to every component that is too large for one thread -- by its size (large_threshold) or because state +
    // coefficients of the one-thread kernel would not fit the register file (it would spill every step)
    const bool spills = std::max((long long)dc.h.fwd.nc() + 4ll * dc.h.fwd.N, dc.h.joint ? (long long)dc.h.adj.nc() + 4ll * dc.h.adj.N : 0ll) *
                                (long long)(sizeof(T) / 4) + 24 > E.one_thread_regs;
    if ((is_large(E, dc.h) || spills) && E.lanes_max >= 8) {
      // lane-block split (lane_split.hpp): the component becomes n_inst * G virtual instances of one lane-local pattern
      // that the register-resident kernels integrate with G threads per trajectory; components it does not suit
      // (dense operators, irregular coupling) stay with the tile kernels / the dense-operator path below
      ClassData split;
      LaneSplitReport rep;
      bool ok = false;
      try {
        // fp32 with an even number of excitations: the class will run two trajectories per thread (packed FFMA2), so the
        // split is sized with every state slot counted twice (it then picks a group twice as wide)
        const bool pair_mode = sizeof(T) == 4 && E.pair_on && E.lane_pair_on && E.E % 2 == 0;
        ok = lane_split(dc.h, (int)sizeof(T), E.lanes_max, E.lane_regs, E.lane_min_eff, split, rep, pair_mode, E.lanes_min);
      } catch (const std::exception& ex) {       // a split that fails its own check is never used: the other paths take over
        rep.why = ex.what();
        ++E.lane_selfcheck_failures;
      }
      if (ok) {
        dc.h = std::move(split);
        ++E.lane_classes;
        E.lane_report = rep;
      } else E.lane_reject = rep.why;
    }
    if (E.unit_on && (dc.h.lanes > 1 || !is_large(E, dc.h))) mark_unit_terms(dc.h);
    dc.n_traj = (long long)dc.h.n_inst * E.E;
    if (dc.n_traj >= (1ll << 31)) throw std::runtime_error("more than 2^31 trajectories in one class");
    if (E.host_only) continue;
    upload<T>(dc.coef, dc.h.coef);
    upload<T>(dc.ic, dc.h.ic);
    upload_int(dc.varmap, dc.h.varmap);
    if (dc.h.joint) {
      upload<T>(dc.coef_adj, dc.h.coef_adj);
      upload<T>(dc.ic_adj, dc.h.ic_adj);
      upload_int(dc.varmap_adj, dc.h.varmap_adj);
    }
    if (dc.h.lanes == 1 && is_large(E, dc.h)) build_large_class<T>(E, dc);
    else resolve_kernels(E, dc);
  }
}

double eps_of(const pasif_engine& E) { return E.dtype == PASIF_F32 ? 1.1920928955078125e-07 : 2.220446049250313e-16; }

void analyze_fwd(pasif_engine& E) {
  E.fwd_dirty = false;                             // allocateDeviceSystemStatesVector: d_fwd_Q <- initial conditions
  auto cds = analyze_forward(E.fwd, eps_of(E), E.dtype == PASIF_F32);
  if (E.dtype == PASIF_F32) build_classes<float>(E, std::move(cds), E.fwd_classes);
  else build_classes<double>(E, std::move(cds), E.fwd_classes);
  E.fwd_ready = true;
}
void analyze_jnt(pasif_engine& E) {
  auto cds = analyze_joint(E.fwd, E.bwd, eps_of(E), E.dtype == PASIF_F32);
  if (E.dtype == PASIF_F32) build_classes<float>(E, std::move(cds), E.joint_classes);
  else build_classes<double>(E, std::move(cds), E.joint_classes);
  E.joint_ready = true;
}

// ---- excitation staging on the device
template <typename T>
void ensure_excitations(pasif_engine& E) {
  const bool half = E.interp_window != 0;
  if (!E.exc_dirty && half == E.exc_half) return;
  if (!E.have_exc) throw std::runtime_error("no excitation set loaded");
  DevBuf& raw = E.d_raw;                           // uploaded by load_exc (stream-ordered before the kernels below)
  if (!half) {
    const long long rows = (long long)std::max(E.L, E.numsteps) + 8;
    E.d_exc.reserve((size_t)rows * E.E * sizeof(T));
    dim3 grid((unsigned)((rows + 31) / 32), (unsigned)((E.E + 31) / 32)), blk(32, 8);
    if (E.raw_f32) prep_exc_full_kernel<T, float><<<grid, blk, 0, E.stream>>>((const float*)raw.p, (T*)E.d_exc.p, (int)E.E, (int)E.L, (int)rows);
    else prep_exc_full_kernel<T, double><<<grid, blk, 0, E.stream>>>((const double*)raw.p, (T*)E.d_exc.p, (int)E.E, (int)E.L, (int)rows);
    E.exc_rows = rows;
  } else {
    const uint32_t W = E.interp_window;
    if (!E.interp_linear && E.interp_table.size() < 2ull * W)
      throw std::runtime_error("interpolation table must hold 2 x window entries (the reference kernel reads IM[W + j], "
                               "src/kernels.cu:124)");
    const int log2W = (int)(unsigned)std::log2f((float)W);
    const long long U = 2ll * E.numsteps + 8;
    DevBuf tab;
    if (E.interp_linear) {
      tab.reserve(2ull * W * sizeof(T));
      gen_linear_table_kernel<T><<<(W + 127) / 128, 128, 0, E.stream>>>((T*)tab.p, (int)W);
    } else upload<T>(tab, E.interp_table);
    E.d_exc.reserve((size_t)U * E.E * sizeof(T));
    const long long n = U * E.E;
    if (E.raw_f32) prep_exc_half_kernel<T, float><<<(unsigned)((n + 255) / 256), 256, 0, E.stream>>>(
        (const float*)raw.p, (const T*)tab.p, (T*)E.d_exc.p, (int)E.E, (int)E.L, U, (int)W, log2W);
    else prep_exc_half_kernel<T, double><<<(unsigned)((n + 255) / 256), 256, 0, E.stream>>>(
        (const double*)raw.p, (const T*)tab.p, (T*)E.d_exc.p, (int)E.E, (int)E.L, U, (int)W, log2W);
    CUDA_CHECK(cudaStreamSynchronize(E.stream));
    tab.release();
    E.exc_rows = U;
  }
  CUDA_CHECK(cudaGetLastError());
  E.exc_dirty = false;
  E.exc_half = half;
}

template <typename T>
void fill_common(const pasif_engine& E, const DevClass& c, pasif::FwdArgs<T>& a) {
  std::memset(&a, 0, sizeof a);
  a.coef = (const T*)c.coef.p;
  a.ic = (const T*)c.ic.p;
  a.exc = (const T*)E.d_exc.p;
  a.n_inst = c.h.n_inst; a.E = (int)E.E; a.n_traj = (int)c.n_traj;
  a.half_mode = E.exc_half ? 1 : 0;
  const T h = (T)(1.0 / (double)E.sample_rate);           // setTimesteps(), reference src/__GpuDriver.cu:632-636
  a.h = h; a.h2 = (T)((double)h / 2.0); a.h6 = (T)((double)h / 6.0);
  a.n_glob = (int)E.fwd.n;
  a.varmap = (const int*)c.varmap.p;
  a.save_every = 1;
}

void begin_timing(pasif_engine& E) {
  E.last_launches = 0;
  E.prof_used = 0;
  E.prof_kinds.clear();
  for (int k = 0; k < 5; ++k) { E.kind_ms[k] = 0; E.kind_launches[k] = 0; }
  CUDA_CHECK(cudaEventRecord(E.ev0, E.stream));
}
void end_timing(pasif_engine& E) {
  CUDA_CHECK(cudaEventRecord(E.ev1, E.stream));
  CUDA_CHECK(cudaStreamSynchronize(E.stream));
  CUDA_CHECK(cudaGetLastError());
  float ms = 0;
  CUDA_CHECK(cudaEventElapsedTime(&ms, E.ev0, E.ev1));
  E.last_ms = ms;
  for (size_t i = 0; i + 1 < E.prof_used; i += 2) {
    float t = 0;
    CUDA_CHECK(cudaEventElapsedTime(&t, E.prof_events[i], E.prof_events[i + 1]));
    E.kind_ms[E.prof_kinds[i / 2]] += t;
  }
}

constexpr size_t kStageSlot = 8u << 20;

// host-side copy / widening split over a few threads: one core moves ~10 GB/s between pageable and pinned memory, which
// would make the staging copy -- not the DMA -- the bottleneck of a large hand-off (C4: 262 MB of excitations per call)
template <typename F>
void host_parallel(size_t n, size_t grain, F&& fn) {          // fn(begin, end) over [0, n)
  unsigned nt = std::min<unsigned>(4, std::max(1u, std::thread::hardware_concurrency() / 2));
  if (n < 2 * grain || nt < 2) { fn((size_t)0, n); return; }
  nt = (unsigned)std::min<size_t>(nt, n / grain);
  const size_t per = (n + nt - 1) / nt;
  std::vector<std::thread> th;
  for (unsigned i = 1; i < nt; ++i) th.emplace_back([&, i] { fn(std::min(n, i * per), std::min(n, (i + 1) * per)); });
  fn((size_t)0, std::min(n, per));
  for (auto& t : th) t.join();
}

double wall_ms() {
  timespec ts;
  clock_gettime(CLOCK_MONOTONIC, &ts);
  return ts.tv_sec * 1e3 + ts.tv_nsec * 1e-6;
}

// host (pageable) -> device through the two pinned slots, stream-ordered on E.stream; returns without waiting for the
// last DMA (consumers are launched on the same stream)
void h2d_staged(pasif_engine& E, void* dst, const void* src, size_t bytes) {
  const double t0 = wall_ms();
  E.h_stage.reserve(2 * kStageSlot);
  size_t i = 0;
  for (size_t off = 0; off < bytes; off += kStageSlot, ++i) {
    const size_t len = std::min(kStageSlot, bytes - off);
    const int slot = (int)(i & 1);
    char* h = (char*)E.h_stage.p + slot * kStageSlot;
    CUDA_CHECK(cudaEventSynchronize(E.ev_stage[slot]));          // the DMA that last read this slot has finished
    const char* sp = (const char*)src + off;
    host_parallel(len, (size_t)1 << 20, [&](size_t b, size_t e) { std::memcpy(h + b, sp + b, e - b); });
    CUDA_CHECK(cudaMemcpyAsync((char*)dst + off, h, len, cudaMemcpyHostToDevice, E.stream));
    CUDA_CHECK(cudaEventRecord(E.ev_stage[slot], E.stream));
  }
  E.last_h2d_ms = wall_ms() - t0;
}

// the same with the samples narrowed to float on the host while they are copied into the pinned slot: half the bytes cross
// the host link.  An fp32 engine rounds every sample to float anyway ((T)raw in the prep kernels), so for it this is
// bit-identical; an fp64 engine does it only on request (option exc_storage=f32).
void h2d_staged_f32(pasif_engine& E, float* dst, const double* src, size_t n) {
  const double t0 = wall_ms();
  E.h_stage.reserve(2 * kStageSlot);
  const size_t per = kStageSlot / sizeof(float);
  size_t i = 0;
  for (size_t off = 0; off < n; off += per, ++i) {
    const size_t len = std::min(per, n - off);
    const int slot = (int)(i & 1);
    float* h = (float*)((char*)E.h_stage.p + slot * kStageSlot);
    CUDA_CHECK(cudaEventSynchronize(E.ev_stage[slot]));
    const double* sp = src + off;
    host_parallel(len, (size_t)1 << 18, [&](size_t b, size_t e) { for (size_t k = b; k < e; ++k) h[k] = (float)sp[k]; });
    CUDA_CHECK(cudaMemcpyAsync(dst + off, h, len * sizeof(float), cudaMemcpyHostToDevice, E.stream));
    CUDA_CHECK(cudaEventRecord(E.ev_stage[slot], E.stream));
  }
  E.last_h2d_ms = wall_ms() - t0;
}

// device -> host (pageable doubles): DMA of chunk i+1 into one pinned slot overlaps the widening copy of chunk i out of
// the other.  The reference returns `reel` values widened to Python floats (src/PASIfgpu.cpp:100-104); so does this.
template <typename T>
void download(pasif_engine& E, const void* dev, double* out, uint64_t n) {
  if (!out || n == 0) return;
  const double t0 = wall_ms();
  E.h_stage.reserve(2 * kStageSlot);
  const uint64_t per = kStageSlot / sizeof(T);
  const uint64_t chunks = (n + per - 1) / per;
  for (int s = 0; s < 2; ++s) CUDA_CHECK(cudaEventSynchronize(E.ev_stage[s]));
  auto issue = [&](uint64_t c) {
    const uint64_t len = std::min(per, n - c * per);
    const int slot = (int)(c & 1);
    CUDA_CHECK(cudaMemcpyAsync((char*)E.h_stage.p + slot * kStageSlot, (const T*)dev + c * per, len * sizeof(T),
                               cudaMemcpyDeviceToHost, E.stream));
    CUDA_CHECK(cudaEventRecord(E.ev_stage[slot], E.stream));
  };
  issue(0);
  for (uint64_t c = 0; c < chunks; ++c) {
    if (c + 1 < chunks) issue(c + 1);
    const int slot = (int)(c & 1);
    CUDA_CHECK(cudaEventSynchronize(E.ev_stage[slot]));
    const T* h = (const T*)((const char*)E.h_stage.p + slot * kStageSlot);
    const uint64_t len = std::min(per, n - c * per);
    double* o = out + c * per;
    host_parallel((size_t)len, (size_t)1 << 18, [&](size_t b, size_t e) {
      if (sizeof(T) == sizeof(double)) std::memcpy(o + b, h + b, (e - b) * sizeof(double));
      else for (size_t i = b; i < e; ++i) o[i] = (double)h[i];
    });
  }
  E.last_d2h_ms = wall_ms() - t0;
}

// class-layout copy [N][n_traj] of the forward state a previous gradient call left behind (E.d_fwd_cur, user layout);
// stream-ordered, so consecutive classes may share the scratch buffer
template <typename T>
const T* gather_start_state(pasif_engine& E, const DevClass& c) {
  const long long tot = (long long)c.h.fwd.N * c.n_traj;
  E.d_q0.reserve((size_t)std::max<long long>(tot, 1) * sizeof(T));
  gather_user_kernel<T><<<(unsigned)((tot + 255) / 256), 256, 0, E.stream>>>(
      (const T*)E.d_fwd_cur.p, (const int*)c.varmap.p, (const T*)c.ic.p, (T*)E.d_q0.p, c.h.fwd.N, c.h.n_inst, (int)E.E,
      (int)c.n_traj, (int)E.fwd.n, c.h.lanes, use_pairs(E, c) ? 2 : 1);
  ++E.last_launches;
  return (const T*)E.d_q0.p;
}

// ---- getAmplitudes / getTrajectory
template <typename T>
void solve_forward(pasif_engine& E, uint32_t save_steps, double* out, uint64_t n_out, bool trajectory) {
  E.use_device();
  if (!E.fwd_ready) throw std::runtime_error("the system must be allocated (pasif_alloc_system) before solving");
  ensure_excitations<T>(E);
  const uint64_t width = (uint64_t)E.fwd.n * E.E;
  uint64_t n_slots = 0;
  if (trajectory) {
    if (save_steps == 0) throw std::runtime_error("saveSteps must be >= 1");
    n_slots = width ? (width * E.numsteps) / save_steps / width : 0;
    if (out && n_out != n_slots * width) throw std::runtime_error("trajectory output buffer has the wrong size");
    E.d_traj.reserve(std::max<uint64_t>(n_slots * width, 1) * sizeof(T));
  } else {
    if (out && n_out != width) throw std::runtime_error("amplitude output buffer must hold N * E values");
    E.d_user.reserve(std::max<uint64_t>(width, 1) * sizeof(T));
  }
  begin_timing(E);
  for (DevClass& c : E.fwd_classes) {
    if (c.large) { large_solve_forward<T>(E, c, trajectory, save_steps, n_slots); continue; }
    pasif::FwdArgs<T> a;
    fill_common<T>(E, c, a);
    a.t0 = 0; a.t1 = (int)E.numsteps;
    if (E.fwd_dirty) a.q_in = gather_start_state<T>(E, c);      // quirk A7: the sweep starts from the current forward state
    if (trajectory) {
      a.save_ptr = (T*)E.d_traj.p; a.save_every = (int)save_steps; a.save_phase = 0; a.save_layout = 1;
      a.save_slots = (int)n_slots;
    } else {
      a.final_user = (T*)E.d_user.p;
    }
    if (use_pairs(E, c)) { launch(E, c.k_fwd2, c.n_traj / 2, a, 0, nullptr, c.h.lanes); ++E.pair_launches; }
    else launch(E, c.k_fwd, c.n_traj, a, 0, nullptr, c.h.lanes);
  }
  end_timing(E);
  E.fwd_dirty = false;                             // resetStatesVectors after getAmplitudes / getTrajectory (:447, :482)
  if (!trajectory) E.d_user_kind = 1;
  download<T>(E, trajectory ? E.d_traj.p : E.d_user.p, out, trajectory ? n_slots * width : width);
}

// ---- getGradient
struct Schedule {
  long long nsp, c, last;
};
template <typename T>
Schedule reference_schedule(uint32_t n) {       // reference src/__GpuDriver.cu:506-508, evaluated in `reel` arithmetic
  Schedule s;
  s.nsp = (long long)(unsigned)std::sqrt((double)n);
  s.c = (long long)(unsigned)std::ceil((T)n / (T)(unsigned)s.nsp);
  s.last = s.c - (s.nsp * s.c) % n;
  return s;
}

typedef pasif_engine::StageGraphT StageGraph;
#include "large_host.inc"
#include "stage_host.inc"
void LargeClassDeleter::operator()(LargeClass* p) const {
  if (p) { p->release(); delete p; }
}

template <typename T>
void solve_gradient(pasif_engine& E, uint32_t save, T* d_user_out) {
  const long long n = E.numsteps;
  const int threads = 256;
  for (DevClass& c : E.joint_classes) {
    if (c.large) { large_solve_gradient<T>(E, c, save, d_user_out); continue; }
    const long long P = c.n_traj;
    const int NF = c.h.fwd.N, NA = c.h.adj.N, NV = (int)c.h.adj.nv.size();
    // two trajectories per thread (fp32): half as many threads, pair kernels on both sides (they share the chunk-trajectory layout)
    const bool pairs = use_pairs(E, c);
    const KernelHandle& k_gfwd = pairs ? c.k_gfwd2 : c.k_gfwd;
    const KernelHandle& k_bwd = pairs ? c.k_bwd2 : c.k_bwd;
    const long long PT = pairs ? P / 2 : P;        // threads
    if (pairs) ++E.pair_launches;
    E.d_lam.reserve((size_t)NA * P * sizeof(T));
    // call-order state (quirk A7), literal schedule only: the adjoint state continues from what the previous call left in
    // the initial-condition buffer, the forward sweep from the current forward state
    const bool carry = E.call_order_ref && !E.compat_clean;
    if (carry && c.lam_dirty && c.lam_carry.cap >= (size_t)NA * P * sizeof(T)) {
      CUDA_CHECK(cudaMemcpyAsync(E.d_lam.p, c.lam_carry.p, (size_t)NA * P * sizeof(T), cudaMemcpyDeviceToDevice, E.stream));
    } else {
      const long long tot = (long long)NA * P;
      init_state_kernel<T><<<(unsigned)((tot + threads - 1) / threads), threads, 0, E.stream>>>(
          (const T*)c.ic_adj.p, (T*)E.d_lam.p, NA, c.h.n_inst, (int)E.E, (int)P, c.h.lanes, pairs ? 2 : 1);
      ++E.last_launches;
    }
    const T* q_start = (carry && E.fwd_dirty) ? gather_start_state<T>(E, c) : nullptr;
    T* q_end = nullptr;                              // class-layout forward state this call leaves behind
    if (carry) {
      E.d_qend.reserve(std::max<size_t>((size_t)NF * P, 1) * sizeof(T));
      q_end = (T*)E.d_qend.p;
    }
    if (n > 0) {
      long long nsp, chunk, S_global = -1;
      if (!E.compat_clean) {
        Schedule s = reference_schedule<T>((uint32_t)n);
        nsp = s.nsp; chunk = s.c;
        S_global = (nsp - 1) + (s.last - 1);       // first captured bwd_setpoint (src/__GpuDriver.cu:541, 727-736)
      } else {
        chunk = (long long)std::ceil(std::sqrt((double)n));
        nsp = (n + chunk - 1) / chunk;
      }
      const size_t slot_elems = (size_t)NF * P;
      E.d_ckpt.reserve((size_t)(nsp + 1) * slot_elems * sizeof(T));
      E.d_dense.reserve(std::max<size_t>((size_t)chunk * NV * P, 1) * sizeof(T));
      T* ckpt = (T*)E.d_ckpt.p;
      // Schedule sized to HBM (exact schedule only): the chunk trajectories of the LAST chunks -- the first ones the
      // adjoint sweep consumes -- are written by the forward sweep itself into whatever device memory is free, so those
      // chunks need no recompute launch.  In the literal reference schedule the replayed chunk is not the swept
      // trajectory (quirk A1: replay starts one sample late), so there every chunk has to be recomputed.
      long long sp_store = nsp;                    // chunks [sp_store, nsp) are stored by the sweep
      // automatic sizing keeps chunks only when the batch saturates the GPU (>= 8 warps per SM): a small batch is
      // latency-bound, its recompute already runs for free beside the adjoint sweep on the second stream, and reading the
      // kept trajectory from HBM instead of the L2-resident chunk buffer would slow it down (C2: 57 ms vs 42 ms)
      const bool tail_wanted = E.tail_budget_gb > 0 || (E.tail_budget_gb < 0 && P >= 256ll * E.n_sm);
      if (E.compat_clean && tail_wanted) {
        const bool ovl_ = E.overlap && !E.profile;
        if (ovl_) E.d_dense2.reserve(std::max<size_t>((size_t)chunk * NV * P, 1) * sizeof(T));   // count it before sizing
        size_t free_b = 0, total_b = 0;
        CUDA_CHECK(cudaMemGetInfo(&free_b, &total_b));
        // automatic budget: 80 % of what is free, and never the last 6 GB -- torch / NCCL in the same process (bench.py,
        // sharded.py) and the engine's own later buffers allocate from the same device; whatever is held here is also
        // given back when another allocation of this engine fails (DevBuf::reserve) or on option trim
        const double avail = (double)(free_b + E.d_tail.cap);
        double budget = std::max(0.0, std::min(avail * 0.8, avail - 6e9));
        if (E.tail_budget_gb > 0) budget = std::min(avail * 0.9, E.tail_budget_gb * 1e9);
        const double chunk_bytes = (double)chunk * NV * P * sizeof(T);
        const long long fit = (long long)(budget / chunk_bytes);
        sp_store = std::max<long long>(0, nsp - fit);
        if (sp_store < nsp) E.d_tail.reserve((size_t)(n - sp_store * chunk) * NV * P * sizeof(T));
      }
      E.last_stored_chunks = nsp - sp_store;
      E.last_chunks = nsp;
      // forward sweep with checkpoints
      {
        pasif::FwdArgs<T> a;
        fill_common<T>(E, c, a);
        a.t0 = 0; a.t1 = (int)n;
        a.save_ptr = ckpt; a.save_every = (int)chunk; a.save_layout = 0; a.save_slots = (int)nsp;
        if (!E.compat_clean) {
          a.save_phase = 0;                        // slot j <- state AFTER step j*c (quirk A1)
          a.q_in = q_start;                        // nullptr: initial conditions
          a.q_out = q_end;                         // what stays in d_fwd_Q if no chunk is recomputed afterwards
          launch(E, k_gfwd, PT, a, 1, nullptr, c.h.lanes);
        } else {
          a.save_phase = 1;                        // slot j <- state BEFORE step j*c; slot 0 = initial condition
          const long long tot = (long long)NF * P;
          init_state_kernel<T><<<(unsigned)((tot + threads - 1) / threads), threads, 0, E.stream>>>(
              (const T*)c.ic.p, ckpt, NF, c.h.n_inst, (int)E.E, (int)P, c.h.lanes, pairs ? 2 : 1);
          ++E.last_launches;
          const long long t_split = std::min<long long>(n, sp_store * chunk);
          if (t_split > 0) {                       // checkpoints only
            a.t1 = (int)t_split;
            launch(E, k_gfwd, PT, a, 1, nullptr, c.h.lanes);
          }
          if (t_split < n) {                       // checkpoints + chunk trajectories, continuing from slot sp_store
            a.q_in = ckpt + (size_t)sp_store * slot_elems;
            a.t0 = (int)t_split; a.t1 = (int)n;
            a.dense_ptr = (T*)E.d_tail.p;
            launch(E, k_gfwd, PT, a, 1, nullptr, c.h.lanes);
          }
        }
      }
      const long long first = E.compat_clean ? 0 : 1;      // the reference never back-propagates chunk 0 (quirk A2)
      // Two streams, two chunk-trajectory buffers: the recompute of chunk k-1 (stream) runs while the adjoint sweep of
      // chunk k (stream2) is still in flight, so the partial last wave of one kernel is filled by CTAs of the other.
      // profile=on serialises them again (per-kernel CUDA-event times are only meaningful without co-running kernels).
      const bool ovl = E.overlap && !E.profile;
      if (ovl) E.d_dense2.reserve(std::max<size_t>((size_t)chunk * NV * P, 1) * sizeof(T));
      T* dense_buf[2] = {(T*)E.d_dense.p, ovl ? (T*)E.d_dense2.p : (T*)E.d_dense.p};
      cudaStream_t sB = ovl ? E.stream2 : E.stream;
      long long it = 0;
      int last_buf = -1;
      for (long long sp = nsp - 1; sp >= first; --sp) {
        const long long start = sp * chunk;
        const long long end = std::min<long long>(n, start + chunk);
        const long long len = end - start;
        if (len <= 0) continue;
        const int buf = (int)(it & 1);
        const bool stored = sp >= sp_store;        // chunk trajectory already in d_tail (written by the sweep)
        pasif::FwdArgs<T> a;
        fill_common<T>(E, c, a);
        if (!stored) {
          if (ovl && it >= 2) CUDA_CHECK(cudaStreamWaitEvent(E.stream, E.ev_bwd[buf], 0));   // buffer consumed
          a.q_in = ckpt + (size_t)sp * slot_elems;
          a.t0 = (int)start; a.t1 = (int)end;
          a.dense_ptr = dense_buf[buf];
          // the last chunk this call recomputes leaves its end state in d_fwd_Q (src/__GpuDriver.cu:565)
          if (q_end && (sp == first || sp == (long long)save)) a.q_out = q_end;
          launch(E, k_gfwd, PT, a, 1, E.stream, c.h.lanes);
        }
        if (ovl) {                                 // (for a stored chunk: orders the adjoint stream after the sweep)
          CUDA_CHECK(cudaEventRecord(E.ev_rec[buf], E.stream));
          CUDA_CHECK(cudaStreamWaitEvent(sB, E.ev_rec[buf], 0));
        }
        pasif::BwdArgs<T> b;
        std::memset(&b, 0, sizeof b);
        b.coef = (const T*)c.coef_adj.p;
        b.lam = (T*)E.d_lam.p;
        b.dense = stored ? (const T*)E.d_tail.p + (size_t)(start - sp_store * chunk) * NV * P : dense_buf[buf];
        // state standing in for index -1: the reference reads the PREVIOUS chunk's checkpoint slot (quirk A4);
        // the exact sweep uses this chunk's own start state
        b.prev = ckpt + (size_t)(E.compat_clean ? sp : sp - 1) * slot_elems;
        b.len = (int)len;
        b.alias_local = E.compat_clean ? -1000000 : (int)(S_global - sp);
        b.n_inst = c.h.n_inst; b.E = (int)E.E; b.n_traj = (int)P;
        b.h = a.h; b.h2 = a.h2; b.h6 = a.h6;
        launch(E, k_bwd, PT, b, 2, sB, c.h.lanes);
        if (ovl) CUDA_CHECK(cudaEventRecord(E.ev_bwd[buf], sB));
        last_buf = buf;
        ++it;
        if (!E.compat_clean && sp == (long long)save) break;
      }
      if (ovl && last_buf >= 0) CUDA_CHECK(cudaStreamWaitEvent(E.stream, E.ev_bwd[last_buf], 0));
    }
    const long long tot = (long long)NA * P;
    scatter_user_kernel<T><<<(unsigned)((tot + threads - 1) / threads), threads, 0, E.stream>>>(
        (const T*)E.d_lam.p, (const int*)c.varmap_adj.p, d_user_out, NA, c.h.n_inst, (int)E.E, (int)P, (int)E.bwd.n, c.h.lanes, pairs ? 2 : 1);
    ++E.last_launches;
    if (carry) {
      // d_bwd_QinitCond <- running adjoint state (src/__GpuDriver.cu:743); d_fwd_Q keeps the last recomputed state
      c.lam_carry.reserve((size_t)NA * P * sizeof(T));
      CUDA_CHECK(cudaMemcpyAsync(c.lam_carry.p, E.d_lam.p, (size_t)NA * P * sizeof(T), cudaMemcpyDeviceToDevice, E.stream));
      c.lam_dirty = true;
      const long long ftot = (long long)NF * P;
      scatter_user_kernel<T><<<(unsigned)((ftot + threads - 1) / threads), threads, 0, E.stream>>>(
          (const T*)q_end, (const int*)c.varmap.p, (T*)E.d_fwd_cur.p, NF, c.h.n_inst, (int)E.E, (int)P, (int)E.fwd.n, c.h.lanes, pairs ? 2 : 1);
      ++E.last_launches;
    }
  }
}

template <typename T>
void gradient_entry(pasif_engine& E, uint32_t save, double* out, uint64_t n_out, bool reduced, void* dev_out) {
  E.use_device();
  if (!E.fwd_ready || !E.joint_ready)
    throw std::runtime_error("the system and the adjoint must be allocated before computing the gradient");
  ensure_excitations<T>(E);
  const uint64_t width = (uint64_t)E.bwd.n * E.E;
  if (!reduced && out && n_out != width) throw std::runtime_error("gradient output buffer must hold N_adj * E values");
  if (reduced && (out || dev_out) && n_out != E.bwd.n) throw std::runtime_error("reduced gradient buffer must hold N_adj values");
  E.d_user.reserve(std::max<uint64_t>(width, 1) * sizeof(T));
  begin_timing(E);
  const bool carry = E.call_order_ref && !E.compat_clean;
  if (carry) {
    // forward state in the user layout: variables of components no adjoint variable depends on are not integrated by the
    // gradient path and keep their initial conditions
    const uint64_t fw = (uint64_t)E.fwd.n * E.E;
    E.d_fwd_cur.reserve(std::max<uint64_t>(fw, 1) * sizeof(T));
    if (!E.fwd_dirty)
      for (DevClass& c : E.fwd_classes) {
        if (c.large) continue;
        const long long tot = (long long)c.h.fwd.N * c.n_traj;
        E.d_q0.reserve((size_t)std::max<long long>(tot, 1) * sizeof(T));
        init_state_kernel<T><<<(unsigned)((tot + 255) / 256), 256, 0, E.stream>>>((const T*)c.ic.p, (T*)E.d_q0.p, c.h.fwd.N, c.h.n_inst,
                                                                                   (int)E.E, (int)c.n_traj, c.h.lanes, 1);
        scatter_user_kernel<T><<<(unsigned)((tot + 255) / 256), 256, 0, E.stream>>>(
            (const T*)E.d_q0.p, (const int*)c.varmap.p, (T*)E.d_fwd_cur.p, c.h.fwd.N, c.h.n_inst, (int)E.E, (int)c.n_traj, (int)E.fwd.n, c.h.lanes, 1);
        E.last_launches += 2;
      }
  }
  solve_gradient<T>(E, save, (T*)E.d_user.p);
  if (carry) E.fwd_dirty = true;
  T* red = nullptr;
  if (reduced) {
    E.d_red.reserve(std::max<uint64_t>(E.bwd.n, 1) * sizeof(T));
    red = dev_out ? (T*)dev_out : (T*)E.d_red.p;
    reduce_excitations_kernel<T><<<(unsigned)((E.bwd.n + 31) / 32), dim3(32, 32), 0, E.stream>>>(
        (const T*)E.d_user.p, red, (int)E.E, (int)E.bwd.n);
    ++E.last_launches;
  }
  end_timing(E);
  E.d_user_kind = 2;
  if (reduced) { if (!dev_out) download<T>(E, red, out, E.bwd.n); }
  else download<T>(E, E.d_user.p, out, width);
}

// sum over excitations of the adjoint state the last pasif_get_gradient left on the device, written to device memory
template <typename T>
void reduce_last(pasif_engine& E, void* dev_out, uint64_t n_out) {
  E.use_device();
  if (E.d_user_kind != 2) throw std::runtime_error("no gradient on the device: call pasif_get_gradient first");
  if (!dev_out) throw std::runtime_error("device_out is NULL");
  if (n_out != E.bwd.n) throw std::runtime_error("reduced gradient buffer must hold N_adj values");
  reduce_excitations_kernel<T><<<(unsigned)((E.bwd.n + 31) / 32), dim3(32, 32), 0, E.stream>>>(
      (const T*)E.d_user.p, (T*)dev_out, (int)E.E, (int)E.bwd.n);
  CUDA_CHECK(cudaGetLastError());
  CUDA_CHECK(cudaStreamSynchronize(E.stream));
}

void set_tensor(Tensor& t, int order, const uint32_t* dims, const double* values, const uint32_t* indices, uint32_t nnz) {
  t.order = order;
  t.dims.assign(dims, dims + order);
  t.val.assign(values, values + nnz);
  t.idx.assign(indices, indices + (size_t)nnz * order);
}
void set_csr(SystemData& s, uint32_t n_rows, uint32_t n_cols, const double* values, const uint32_t* indices,
             const uint32_t* indptr, uint32_t nnz) {
  if (n_rows != n_cols) throw std::runtime_error("K must be square");
  s.n = n_rows;
  s.kval.assign(values, values + nnz);
  s.kidx.assign(indices, indices + nnz);
  s.kptr.assign(indptr, indptr + n_rows + 1);
  s.have_k = true;
}

void load_exc(pasif_engine& E, const double* exc, uint32_t n_exc, uint32_t length, uint32_t sample_rate) {
  if (n_exc != E.E || length != E.L || sample_rate != E.sample_rate) ++E.cfg_gen;   // (same shape: the device buffers stay where they are)
  if (!exc || n_exc == 0 || length == 0) throw std::runtime_error("excitation set must be a non-empty [E][L] array");
  if (sample_rate == 0) throw std::runtime_error("sampleRate must be > 0");
  if (E.have_exc && n_exc != E.E && (E.fwd_ready || E.joint_ready)) {
    // trajectory count changes: classes keep per-instance tables only, so nothing to rebuild except n_traj
    for (auto* v : {&E.fwd_classes, &E.joint_classes})
      for (auto& c : *v) c.n_traj = (long long)c.h.n_inst * n_exc;
  }
  if (E.have_exc && (n_exc != E.E || length != E.L)) E.d_tail.release();   // sized for the previous batch
  if (E.have_exc && n_exc != E.E) {                // a different batch: nothing a previous call left behind applies to it
    E.fwd_dirty = false;
    for (auto& c : E.joint_classes) c.lam_dirty = false;
  }
  E.E = n_exc; E.L = length; E.sample_rate = sample_rate;
  if (!E.host_only) {
    // straight to the device, as the reference does (src/__GpuDriver.cu:343-381), but through pinned staging and
    // without a host-side copy; the layout/precision pass runs on the device before the next solve
    CUDA_CHECK(cudaSetDevice(E.device));
    const size_t n = (size_t)n_exc * length;
    E.raw_f32 = E.dtype == PASIF_F32 || E.exc_storage_f32;
    E.last_h2d_bytes = n * (E.raw_f32 ? sizeof(float) : sizeof(double));
    E.d_raw.reserve(E.last_h2d_bytes);
    if (E.raw_f32) h2d_staged_f32(E, (float*)E.d_raw.p, exc, n);
    else h2d_staged(E, E.d_raw.p, exc, n * sizeof(double));
  }
  E.have_exc = true;
  E.exc_dirty = true;
}

}  // namespace

// ====================================================================================== C ABI
#define PASIF_TRY(E_) \
  if (!(E_)) return 1; \
  try {
#define PASIF_CATCH(E_)                \
  }                                    \
  catch (const std::exception& ex) {   \
    (E_)->err = ex.what();             \
    return 2;                          \
  }                                    \
  catch (...) {                        \
    (E_)->err = "unknown error";       \
    return 3;                          \
  }                                    \
  return 0;

extern "C" {

int pasif_create(pasif_engine** out, const double* excitations, uint32_t n_excitations, uint32_t length,
                 uint32_t sample_rate, uint32_t numsteps, int dtype, int device) {
  if (!out) return 1;
  *out = nullptr;
  std::unique_ptr<pasif_engine> e(new pasif_engine);
  try {
    if (dtype != PASIF_F64 && dtype != PASIF_F32) throw std::runtime_error("dtype must be PASIF_F64 or PASIF_F32");
    e->dtype = dtype;
    e->host_only = device < 0;
    e->device = device < 0 ? 0 : device;
    e->numsteps = numsteps;
    if (!e->host_only) {
      int count = 0;
      cudaError_t st = cudaGetDeviceCount(&count);
      if (st != cudaSuccess || count == 0)
        throw std::runtime_error(std::string("no CUDA device available (") + cudaGetErrorString(st) +
                                 "): pasif_b200 has no CPU fallback");
      if (device >= count) throw std::runtime_error("GPU id " + std::to_string(device) + " out of range");
      CUDA_CHECK(cudaSetDevice(device));
      CUDA_CHECK(cudaDeviceGetAttribute(&e->n_sm, cudaDevAttrMultiProcessorCount, device));
      CUDA_CHECK(cudaStreamCreateWithFlags(&e->stream, cudaStreamNonBlocking));
      CUDA_CHECK(cudaStreamCreateWithFlags(&e->stream2, cudaStreamNonBlocking));
      for (int i = 0; i < 2; ++i) {
        CUDA_CHECK(cudaEventCreateWithFlags(&e->ev_rec[i], cudaEventDisableTiming));
        CUDA_CHECK(cudaEventCreateWithFlags(&e->ev_bwd[i], cudaEventDisableTiming));
      }
      for (int i = 0; i < 2; ++i) CUDA_CHECK(cudaEventCreateWithFlags(&e->ev_stage[i], cudaEventDisableTiming));
      CUDA_CHECK(cudaEventCreate(&e->ev0));
      CUDA_CHECK(cudaEventCreate(&e->ev1));
    }
    load_exc(*e, excitations, n_excitations, length, sample_rate);
  } catch (const std::exception& ex) {
    g_create_error = ex.what();
    return 2;
  }
  *out = e.release();
  return 0;
}

void pasif_destroy(pasif_engine* e) {
  if (!e) return;
  if (!e->host_only) {
    cudaSetDevice(e->device);
    free_classes(e->fwd_classes);
    free_classes(e->joint_classes);
    for (DevBuf* b : {&e->d_tail, &e->d_raw, &e->d_exc, &e->d_user, &e->d_traj, &e->d_ckpt, &e->d_dense, &e->d_dense2, &e->d_lam, &e->d_red,
                      &e->d_fwd_cur, &e->d_q0, &e->d_qend, &e->stage_sched}) b->release();
    for (int i = 0; i < 2; ++i) {
      if (e->ev_rec[i]) cudaEventDestroy(e->ev_rec[i]);
      if (e->ev_bwd[i]) cudaEventDestroy(e->ev_bwd[i]);
    }
    if (e->stream2) cudaStreamDestroy(e->stream2);
    for (StageBufs*& sb : e->stage_bufs)
      if (sb) { sb->release(); delete sb; sb = nullptr; }
    if (e->stage_graph.exec) cudaGraphExecDestroy(e->stage_graph.exec);
    for (cudaStream_t st : e->part_streams) cudaStreamDestroy(st);
    for (cudaEvent_t ev : e->part_events) cudaEventDestroy(ev);
    if (e->stream) cudaStreamSynchronize(e->stream);
    e->h_stage.release();
    for (int i = 0; i < 2; ++i) if (e->ev_stage[i]) cudaEventDestroy(e->ev_stage[i]);
    for (cudaEvent_t ev : e->prof_events) cudaEventDestroy(ev);
    if (e->ev0) cudaEventDestroy(e->ev0);
    if (e->ev1) cudaEventDestroy(e->ev1);
    if (e->stream) cudaStreamDestroy(e->stream);
  }
  delete e;
}

const char* pasif_last_error(const pasif_engine* e) { return e ? e->err.c_str() : g_create_error.c_str(); }

int pasif_load_excitations(pasif_engine* e, const double* excitations, uint32_t n_excitations, uint32_t length,
                           uint32_t sample_rate) {
  PASIF_TRY(e) load_exc(*e, excitations, n_excitations, length, sample_rate); PASIF_CATCH(e)
}

int pasif_set_fwd_k(pasif_engine* e, uint32_t n_rows, uint32_t n_cols, const double* values, const uint32_t* indices,
                    const uint32_t* indptr, uint32_t nnz) {
  PASIF_TRY(e) set_csr(e->fwd, n_rows, n_cols, values, indices, indptr, nnz); e->fwd_ready = e->joint_ready = false; PASIF_CATCH(e)
}
int pasif_set_fwd_gamma(pasif_engine* e, const uint32_t dims[3], const double* values, const uint32_t* indices, uint32_t nnz) {
  PASIF_TRY(e) set_tensor(e->fwd.G, 3, dims, values, indices, nnz); e->fwd_ready = e->joint_ready = false; PASIF_CATCH(e)
}
int pasif_set_fwd_lambda(pasif_engine* e, const uint32_t dims[4], const double* values, const uint32_t* indices, uint32_t nnz) {
  PASIF_TRY(e) set_tensor(e->fwd.L, 4, dims, values, indices, nnz); e->fwd_ready = e->joint_ready = false; PASIF_CATCH(e)
}
int pasif_set_fwd_psi(pasif_engine* e, const uint32_t dims[5], const double* values, const uint32_t* indices, uint32_t nnz) {
  PASIF_TRY(e) set_tensor(e->fwd.P, 5, dims, values, indices, nnz); e->fwd_ready = e->joint_ready = false; PASIF_CATCH(e)
}
int pasif_set_fwd_force(pasif_engine* e, const double* dense, uint32_t n) {
  PASIF_TRY(e) e->fwd.F.assign(dense, dense + n); e->fwd.have_f = true; e->fwd_ready = e->joint_ready = false; PASIF_CATCH(e)
}
int pasif_set_fwd_ic(pasif_engine* e, const double* ic, uint32_t n) {
  PASIF_TRY(e) e->fwd.IC.assign(ic, ic + n); e->fwd.have_ic = true; e->fwd_ready = e->joint_ready = false; PASIF_CATCH(e)
}
int pasif_alloc_system(pasif_engine* e) {
  PASIF_TRY(e)
  if (!e->host_only) e->use_device();
  analyze_fwd(*e);
  PASIF_CATCH(e)
}

int pasif_set_bwd_k(pasif_engine* e, uint32_t n_rows, uint32_t n_cols, const double* values, const uint32_t* indices,
                    const uint32_t* indptr, uint32_t nnz) {
  PASIF_TRY(e) set_csr(e->bwd, n_rows, n_cols, values, indices, indptr, nnz); e->joint_ready = false; PASIF_CATCH(e)
}
int pasif_set_bwd_gamma(pasif_engine* e, const uint32_t dims[3], const double* values, const uint32_t* indices, uint32_t nnz) {
  PASIF_TRY(e) set_tensor(e->bwd.G, 3, dims, values, indices, nnz); e->joint_ready = false; PASIF_CATCH(e)
}
int pasif_set_bwd_lambda(pasif_engine* e, const uint32_t dims[4], const double* values, const uint32_t* indices, uint32_t nnz) {
  PASIF_TRY(e) set_tensor(e->bwd.L, 4, dims, values, indices, nnz); e->joint_ready = false; PASIF_CATCH(e)
}
int pasif_set_bwd_psi(pasif_engine* e, const uint32_t dims[5], const double* values, const uint32_t* indices, uint32_t nnz) {
  PASIF_TRY(e) set_tensor(e->bwd.P, 5, dims, values, indices, nnz); e->joint_ready = false; PASIF_CATCH(e)
}
int pasif_set_bwd_force(pasif_engine* e, const double* dense, uint32_t n) {
  PASIF_TRY(e) e->bwd.F.assign(dense, dense + n); e->bwd.have_f = true; PASIF_CATCH(e)
}
int pasif_set_bwd_ic(pasif_engine* e, const double* ic, uint32_t n) {
  PASIF_TRY(e) e->bwd.IC.assign(ic, ic + n); e->bwd.have_ic = true; e->joint_ready = false; PASIF_CATCH(e)
}
int pasif_alloc_adjoint(pasif_engine* e) {
  PASIF_TRY(e)
  if (!e->host_only) e->use_device();
  analyze_jnt(*e);
  PASIF_CATCH(e)
}

int pasif_set_interp(pasif_engine* e, const double* table, uint32_t table_len, uint32_t window) {
  PASIF_TRY(e)
  e->interp_linear = false;
  if (window == 0) { e->interp_table.clear(); e->interp_window = 0; }
  else {
    if (!table || table_len == 0) throw std::runtime_error("empty interpolation table");
    e->interp_table.assign(table, table + table_len);
    e->interp_window = window;
  }
  e->exc_dirty = true;
  PASIF_CATCH(e)
}
int pasif_set_interp_linear(pasif_engine* e, uint32_t window) {
  PASIF_TRY(e)
  if (window == 0 || (window & (window - 1)) != 0)
    throw std::runtime_error("the interpolation window must be a power of two (the reference indexes the table with u & (W - 1), "
                             "src/kernels.cu:114)");
  e->interp_table.clear();
  e->interp_window = window;
  e->interp_linear = true;
  e->exc_dirty = true;
  PASIF_CATCH(e)
}
int pasif_set_modulation(pasif_engine* e, const double* buffer, uint32_t n) {
  PASIF_TRY(e) e->modulation.assign(buffer, buffer + n); PASIF_CATCH(e)
}

int pasif_set_option(pasif_engine* e, const char* key, const char* value) {
  PASIF_TRY(e)
  const std::string k(key ? key : ""), v(value ? value : "");
  ++e->cfg_gen;
  if (k == "stage_graph") {
    e->stage_graph_on = (v == "on" || v == "1");
  } else if (k == "compat") {
    if (v == "reference") e->compat_clean = false;
    else if (v == "clean") e->compat_clean = true;
    else throw std::runtime_error("compat must be 'reference' or 'clean'");
  } else if (k == "call_order") {
    // "reference": repeated solver calls behave like the reference driver (quirk A7: _getGradient continues from the forward /
    // adjoint state the previous call left behind; getAmplitudes / getTrajectory reset the forward state afterwards);
    // "independent": every call starts from the initial conditions
    if (v == "reference") e->call_order_ref = true;
    else if (v == "independent") { e->call_order_ref = false; e->fwd_dirty = false; for (auto& c : e->joint_classes) c.lam_dirty = false; }
    else throw std::runtime_error("call_order must be 'reference' or 'independent'");
  } else if (k == "jit") {
    e->jit_on = (v == "on" || v == "1");
  } else if (k == "block") {
    int b = std::atoi(v.c_str());
    if (b < 32 || b > PASIF_BLOCK || (b & 31)) throw std::runtime_error("block must be a multiple of 32 in [32, " + std::to_string(PASIF_BLOCK) + "]");
    e->block = b;
  } else if (k == "gemm") {
    // "on" / "stage": per-stage batched GEMM (stage_kernels.cuh); "tile": GEMM inside the tile kernels; "off": row programs
    e->gemm_on = (v == "on" || v == "1" || v == "stage" || v == "tile");
    e->gemm_stage = (v != "tile");
  } else if (k == "stage_parts") {
    const int n = std::atoi(v.c_str());
    if (n < 1 || n > 16) throw std::runtime_error("stage_parts must be in [1, 16]");
    e->stage_nparts = n;
  } else if (k == "update_u") {
    e->update_u = std::atoi(v.c_str());
  } else if (k == "update_stream") {
    e->update_stream = (v == "on" || v == "1");
  } else if (k == "stream_tt") {
    const int t = std::atoi(v.c_str());
    if (t != 0 && t != 2 && t != 4 && t != 8) throw std::runtime_error("stream_tt must be 0 (auto), 2, 4 or 8");
    e->stream_tt = t;
  } else if (k == "stream_ctas") {
    const int t = std::atoi(v.c_str());
    if (t < 1 || t > 2) throw std::runtime_error("stream_ctas must be 1 or 2");
    e->stream_ctas = t;
  } else if (k == "bwd_priority") {
    // scheduling priority of the adjoint stream against the recompute stream: "high" | "low" | "default"
    if (e->host_only) return 0;
    CUDA_CHECK(cudaSetDevice(e->device));
    int lo = 0, hi = 0;
    CUDA_CHECK(cudaDeviceGetStreamPriorityRange(&lo, &hi));      // lo = least priority (numerically greatest)
    if (e->stream2) { CUDA_CHECK(cudaStreamSynchronize(e->stream2)); CUDA_CHECK(cudaStreamDestroy(e->stream2)); e->stream2 = nullptr; }
    if (v == "high") CUDA_CHECK(cudaStreamCreateWithPriority(&e->stream2, cudaStreamNonBlocking, hi));
    else if (v == "low") CUDA_CHECK(cudaStreamCreateWithPriority(&e->stream2, cudaStreamNonBlocking, lo));
    else CUDA_CHECK(cudaStreamCreateWithFlags(&e->stream2, cudaStreamNonBlocking));
  } else if (k == "stage_direct") {
    e->stage_direct = (v == "on" || v == "1");
  } else if (k == "tc_bk") {
    const int b = std::atoi(v.c_str());
    if (b != 16 && b != 32) throw std::runtime_error("tc_bk must be 16 or 32");
    e->tc_bk = b;
  } else if (k == "tc_bn") {
    e->tc_bn = std::atoi(v.c_str());
  } else if (k == "stream_smem_kb") {
    const int t = std::atoi(v.c_str());
    if (t != 0 && (t < 16 || t > 226)) throw std::runtime_error("stream_smem_kb must be 0 (auto) or 16..226");
    e->stream_smem_kb = t;
  } else if (k == "stream_grid") {
    const int t = std::atoi(v.c_str());
    if (t < 0 || t > 4) throw std::runtime_error("stream_grid must be 0 (auto) .. 4");
    e->stream_grid = t;
  } else if (k == "stream_grid_ctas") {
    const int t = std::atoi(v.c_str());
    if (t < 0 || t > 4096) throw std::runtime_error("stream_grid_ctas must be 0 (auto) .. 4096");
    e->stream_grid_ctas = t;
  } else if (k == "stream_dynamic") {
    e->stream_dynamic = (v == "on" || v == "1");
  } else if (k == "dg_bm") {
    const int t = std::atoi(v.c_str());
    if (t != 32 && t != 64) throw std::runtime_error("dg_bm must be 32 or 64");
    e->dg_bm = t;
  } else if (k == "tc_stages") {
    const int t = std::atoi(v.c_str());
    if (t != 0 && (t < 2 || t > 6)) throw std::runtime_error("tc_stages must be 0 (auto) or 2..6");
    e->tc_stages = t;
  } else if (k == "pdl") {
    g_stage_pdl = (v == "on" || v == "1");              // (process-wide: programmatic dependent launch of the dense-operator kernels)
  } else if (k == "stage_dephase") {
    e->stage_dephase = (v == "on" || v == "1");
  } else if (k == "tc_pair") {
    e->tc_pair = (v == "on" || v == "1");
  } else if (k == "tc_pair_nt") {
    const int t = std::atoi(v.c_str());
    if (t != 1 && t != 2) throw std::runtime_error("tc_pair_nt must be 1 or 2");
    e->tc_pair_nt = t;
  } else if (k == "tc_multicast") {
    e->tc_multicast = (v == "on" || v == "1");
  } else if (k == "tc_split") {
    // "fused": the tcgen05 GEMM splits the stage vector in shared memory; "presplit": the update kernel writes hi / lo copies
    if (v != "fused" && v != "presplit") throw std::runtime_error("tc_split must be 'fused' or 'presplit'");
    e->tc_fused_split = (v == "fused");
  } else if (k == "gemm_band") {
    e->gemm_band = (v == "on" || v == "1");
  } else if (k == "tcgen05") {
    e->tcgen05_on = (v == "on" || v == "1");
  } else if (k == "lanes") {
    e->lanes_max = (v == "off" || v == "0") ? 0 : (v == "on" ? 1024 : std::atoi(v.c_str()));
  } else if (k == "unit_terms") {
    e->unit_on = (v == "on" || v == "1");
  } else if (k == "pair") {
    e->pair_on = (v == "on" || v == "1");
  } else if (k == "lane_pair") {
    e->lane_pair_on = (v == "on" || v == "1");
  } else if (k == "pair_regs") {
    e->pair_regs = std::atoi(v.c_str());
  } else if (k == "aot") {
    e->aot_on = (v == "on" || v == "1");
  } else if (k == "jit_minb_fwd") {
    e->jit_minb_fwd = std::atoi(v.c_str());
  } else if (k == "jit_minb_bwd") {
    e->jit_minb_bwd = std::atoi(v.c_str());
  } else if (k == "lane_minb_fwd") {
    e->lane_minb_fwd = std::atoi(v.c_str());
  } else if (k == "lane_minb_bwd") {
    e->lane_minb_bwd = std::atoi(v.c_str());
  } else if (k == "lanes_min") {
    const int t = std::atoi(v.c_str());
    if (t != 2 && t != 4 && t != 8 && t != 16 && t != 32) throw std::runtime_error("lanes_min must be 2, 4, 8, 16 or 32");
    e->lanes_min = t;
  } else if (k == "one_thread_regs") {
    e->one_thread_regs = std::atoi(v.c_str());
  } else if (k == "lane_min_eff") {
    e->lane_min_eff = std::atof(v.c_str());
  } else if (k == "lane_regs") {
    e->lane_regs = std::atoi(v.c_str());
  } else if (k == "large_threshold") {
    // one-thread register kernels hold at most 48 coupled variables; larger components go to the lane-block /
    // dense-operator / tile paths whatever the option says
    const int t = std::atoi(v.c_str());
    if (t < 0 || t > 48) throw std::runtime_error("large_threshold must be in [0, 48]");
    e->large_threshold = t;
  } else if (k == "overlap") {
    e->overlap = (v == "on" || v == "1");
  } else if (k == "profile") {
    e->profile = (v == "on" || v == "1");
  } else if (k == "exc_storage") {
    // fp64 engines: "f32" narrows the excitation samples to float on the host before they cross the link and keeps them
    // so on the device (half the upload and half the footprint of the staged set; every sample carries a relative
    // rounding error <= 6e-8, measured effect on the C2 gradient in tests/test_excitations_gpu.py); "f64": as handed over.
    // fp32 engines always stage in float (bit-identical).  Takes effect at the next pasif_load_excitations.
    if (v != "f32" && v != "f64") throw std::runtime_error("exc_storage must be 'f32' or 'f64'");
    e->exc_storage_f32 = (v == "f32");
  } else if (k == "trim") {
    // gives back device memory held between calls: the kept chunk trajectories and the work arrays of the solvers
    if (!e->host_only) {
      CUDA_CHECK(cudaSetDevice(e->device));
      for (DevBuf* b : {&e->d_tail, &e->d_traj, &e->d_ckpt, &e->d_dense, &e->d_dense2}) b->release();
    }
  } else if (k == "tail_budget_gb") {
    e->tail_budget_gb = (v == "auto") ? -1.0 : std::atof(v.c_str());
  } else if (k == "numsteps") {
    e->numsteps = (uint32_t)std::strtoul(v.c_str(), nullptr, 10);
    e->exc_dirty = true;
  } else throw std::runtime_error("unknown option '" + k + "'");
  PASIF_CATCH(e)
}

int pasif_get_amplitudes(pasif_engine* e, double* out, uint64_t n_out) {
  PASIF_TRY(e)
  if (e->dtype == PASIF_F32) solve_forward<float>(*e, 1, out, n_out, false);
  else solve_forward<double>(*e, 1, out, n_out, false);
  PASIF_CATCH(e)
}

uint64_t pasif_trajectory_size(const pasif_engine* e, uint32_t save_steps) {
  if (!e || save_steps == 0) return 0;
  const uint64_t width = (uint64_t)e->fwd.n * e->E;
  return width ? (width * e->numsteps) / save_steps / width * width : 0;
}

int pasif_get_trajectory(pasif_engine* e, uint32_t save_steps, double* out, uint64_t n_out) {
  PASIF_TRY(e)
  if (e->dtype == PASIF_F32) solve_forward<float>(*e, save_steps, out, n_out, true);
  else solve_forward<double>(*e, save_steps, out, n_out, true);
  PASIF_CATCH(e)
}

int pasif_get_gradient(pasif_engine* e, uint32_t save, double* out, uint64_t n_out) {
  PASIF_TRY(e)
  if (e->dtype == PASIF_F32) gradient_entry<float>(*e, save, out, n_out, false, nullptr);
  else gradient_entry<double>(*e, save, out, n_out, false, nullptr);
  PASIF_CATCH(e)
}
int pasif_get_gradient_reduced(pasif_engine* e, uint32_t save, double* out, uint64_t n_out) {
  PASIF_TRY(e)
  if (e->dtype == PASIF_F32) gradient_entry<float>(*e, save, out, n_out, true, nullptr);
  else gradient_entry<double>(*e, save, out, n_out, true, nullptr);
  PASIF_CATCH(e)
}
int pasif_get_gradient_reduced_device(pasif_engine* e, uint32_t save, void* device_out, uint64_t n_out) {
  PASIF_TRY(e)
  if (!device_out) throw std::runtime_error("device_out is NULL");
  if (e->dtype == PASIF_F32) gradient_entry<float>(*e, save, nullptr, n_out, true, device_out);
  else gradient_entry<double>(*e, save, nullptr, n_out, true, device_out);
  PASIF_CATCH(e)
}

int pasif_reduce_last_gradient_device(pasif_engine* e, void* device_out, uint64_t n_out) {
  PASIF_TRY(e)
  if (e->dtype == PASIF_F32) reduce_last<float>(*e, device_out, n_out);
  else reduce_last<double>(*e, device_out, n_out);
  PASIF_CATCH(e)
}

int pasif_query(const pasif_engine* e, const char* key, double* value) {
  if (!e || !key || !value) return 1;
  const std::string k(key);
  auto inst = [](const std::vector<DevClass>& v) { double s = 0; for (auto& c : v) s += c.h.n_inst; return s; };
  if (k == "n_fwd") *value = e->fwd.n;
  else if (k == "n_adj") *value = e->bwd.n;
  else if (k == "n_excitations") *value = e->E;
  else if (k == "length") *value = e->L;
  else if (k == "numsteps") *value = e->numsteps;
  else if (k == "fwd_classes") *value = (double)e->fwd_classes.size();
  else if (k == "fwd_instances") *value = inst(e->fwd_classes);
  else if (k == "joint_classes") *value = (double)e->joint_classes.size();
  else if (k == "joint_instances") *value = inst(e->joint_classes);
  else if (k == "last_device_ms") *value = e->last_ms;
  else if (k == "last_launches") *value = (double)e->last_launches;
  else if (k == "stored_chunks") *value = (double)e->last_stored_chunks;
  else if (k == "chunks") *value = (double)e->last_chunks;
  else if (k == "last_h2d_ms") *value = e->last_h2d_ms;
  else if (k == "last_h2d_bytes") *value = (double)e->last_h2d_bytes;
  else if (k == "last_d2h_ms") *value = e->last_d2h_ms;
  else if (k == "jit_compiles") *value = (double)e->jit_compiles;
  else if (k == "jit_cache_hits") *value = (double)e->jit_cache_hits;
  else if (k == "aot_hits") *value = (double)e->aot_hits;
  else if (k == "pair_launches") *value = (double)e->pair_launches;
  else if (k == "dtype") *value = e->dtype;
  else if (k == "large_classes") *value = (double)e->large_classes;
  else if (k == "lane_classes") *value = (double)e->lane_classes;
  else if (k == "lane_selfcheck_failures") *value = (double)e->lane_selfcheck_failures;
  else if (k == "lane_G") *value = e->lane_report.G;
  else if (k == "lane_R") *value = e->lane_report.R;
  else if (k == "lane_RA") *value = e->lane_report.RA;
  else if (k == "lane_nv") *value = e->lane_report.nv;
  else if (k == "lane_nc_fwd") *value = e->lane_report.nc_fwd;
  else if (k == "lane_nc_adj") *value = e->lane_report.nc_adj;
  else if (k == "lane_halo_fwd") *value = e->lane_report.halo_fwd;
  else if (k == "lane_halo_adj") *value = e->lane_report.halo_adj;
  else if (k == "lane_halo_p") *value = e->lane_report.halo_p;
  else if (k == "lane_regs_est") *value = e->lane_report.regs_est;
  else if (k == "lane_efficiency") *value = e->lane_report.efficiency;
  else if (k == "large_gemm_classes") *value = (double)e->large_gemm_classes;
  else if (k == "large_stage_classes") *value = (double)e->large_stage_classes;
  else if (k == "gemm_launches") *value = (double)e->gemm_launches;
  else if (k == "stage_enqueue_us") *value = (double)e->stage_enqueue_us;
  else if (k == "stage_graph_replays") *value = (double)e->stage_graph_replays;
  else if (k == "stage_direct") *value = stage_direct(*e, e->dtype == PASIF_F32 ? 4 : 8) ? 1.0 : 0.0;
  else if (k.rfind("stage_band_", 0) == 0) {
    // rows / columns of K' that go through the per-stage GEMM of the (first) dense-operator class:
    // stage_band_rows_fwd | stage_band_rows_adj | stage_band_cols_fwd | stage_band_cols_adj; stage_band_nr_fwd | _adj:
    // padded rows of the state arrays
    *value = 0;
    const bool fwd = k.size() > 4 && k.compare(k.size() - 4, 4, "_fwd") == 0;
    const bool rows = k.find("_rows_") != std::string::npos;
    for (auto* v : {&e->joint_classes, &e->fwd_classes})
      for (auto& c : *v)
        if (*value == 0 && c.large && c.large->stage) {
          const LargeSide& sd = fwd ? c.large->fwd : c.large->adj;
          if (k.find("_nr_") != std::string::npos) *value = (double)sd.tb.NR;
          else *value = rows ? (double)(sd.band_hi - sd.band_lo) : (double)(sd.kband_hi - sd.kband_lo);
        }
  }
  else if (k == "aot_table_size") *value = (double)aot_table().size();
  else if (k == "ms_fwd") *value = e->kind_ms[0];
  else if (k == "ms_gfwd") *value = e->kind_ms[1];
  else if (k == "ms_bwd") *value = e->kind_ms[2];
  else if (k == "launches_fwd") *value = (double)e->kind_launches[0];
  else if (k == "launches_gfwd") *value = (double)e->kind_launches[1];
  else if (k == "launches_bwd") *value = (double)e->kind_launches[2];
  else if (k == "ms_upd_fwd") *value = e->kind_ms[3];
  else if (k == "ms_upd_adj") *value = e->kind_ms[4];
  else if (k == "launches_upd_fwd") *value = (double)e->kind_launches[3];
  else if (k == "launches_upd_adj") *value = (double)e->kind_launches[4];
  else return 2;
  return 0;
}

int pasif_emit_patterns(pasif_engine* e, char* buf, uint64_t cap, uint64_t* needed) {
  PASIF_TRY(e)
  std::string s;
  for (auto* v : {&e->fwd_classes, &e->joint_classes})
    for (auto& c : *v) {
      s += emit_class_traits(c.h);
      s += std::string(c.h.joint ? "PASIF_AOT_JOINT(" : "PASIF_AOT_FWD(") + c.h.name + ", 0x" + detail::hex64(c.h.hash) + "ull)\n";
    }
  if (needed) *needed = s.size() + 1;
  if (buf && cap) {
    const size_t n = std::min<size_t>(s.size(), cap - 1);
    std::memcpy(buf, s.data(), n);
    buf[n] = 0;
  }
  PASIF_CATCH(e)
}

int pasif_debug_jit_log(pasif_engine* e, const char* arch, char* buf, uint64_t cap, uint64_t* needed) {
  PASIF_TRY(e)
  // compiles the run-time specialised kernels of every loaded class for `arch` (e.g. "sm_100a") with ptxas -v and returns
  // the logs; needs NVRTC but no device (host-only engines included)
  g_nvrtc.load();
  std::string out;
  const std::string a = std::string("--gpu-architecture=") + (arch ? arch : "sm_100a");
  for (auto* v : {&e->fwd_classes, &e->joint_classes})
    for (auto& c : *v) {
      if (c.large || (c.h.lanes == 1 && is_large(*e, c.h))) continue;   // tile / dense-operator classes are prebuilt kernels
      const std::string src = jit_source(c.h, e->dtype);
      nvrtcProgram prog = nullptr;
      if (g_nvrtc.CreateProgram(&prog, src.c_str(), "pasif_jit.cu", 0, nullptr, nullptr) != 0) throw std::runtime_error("nvrtcCreateProgram failed");
      const char* opts[] = {a.c_str(), "-std=c++17", "-lineinfo", "--ptxas-options=-v"};
      const int rc = g_nvrtc.CompileProgram(prog, 4, opts);
      if (const char* dump = std::getenv("PASIF_B200_DUMP_CUBIN")) {      // (debug: the cubin beside the log, for cuobjdump -sass)
        size_t cn = 0;
        if (rc == 0 && g_nvrtc.GetCUBINSize(prog, &cn) == 0 && cn) {
          std::vector<char> cb(cn);
          g_nvrtc.GetCUBIN(prog, cb.data());
          std::ofstream(std::string(dump) + "/" + c.h.name + "_lanes" + std::to_string(c.h.lanes) + ".cubin", std::ios::binary).write(cb.data(), (std::streamsize)cn);
        }
      }
      size_t n = 0;
      g_nvrtc.GetProgramLogSize(prog, &n);
      std::string log(n, '\0');
      if (n) g_nvrtc.GetProgramLog(prog, &log[0]);
      while (!log.empty() && log.back() == '\0') log.pop_back();
      g_nvrtc.DestroyProgram(&prog);
      out += "== " + c.h.name + " lanes=" + std::to_string(c.h.lanes) + " rc=" + std::to_string(rc) + "\n" + log + "\n";
    }
  if (needed) *needed = out.size() + 1;
  if (buf && cap) {
    const size_t n = std::min<size_t>(out.size(), cap - 1);
    std::memcpy(buf, out.data(), n);
    buf[n] = 0;
  }
  PASIF_CATCH(e)
}

int pasif_debug_gemm(int device, int dtype, int mode, uint32_t n, uint64_t n_traj, const double* kd, const double* x, double* y) {
  try {
    CUDA_CHECK(cudaSetDevice(device));
    if (dtype == PASIF_F32) debug_gemm<float>(mode, (int)n, (long long)n_traj, kd, x, y);
    else debug_gemm<double>(mode, (int)n, (long long)n_traj, kd, x, y);
  } catch (const std::exception& ex) {
    g_create_error = ex.what();
    return 2;
  }
  return 0;
}

int pasif_fma_peak(int device, int dtype, double* tflops) {
  if (!tflops) return 1;
  try {
    CUDA_CHECK(cudaSetDevice(device));
    cudaDeviceProp prop;
    CUDA_CHECK(cudaGetDeviceProperties(&prop, device));
    const int threads = 256, blocks = prop.multiProcessorCount * 8, iters = 1 << 12;   // x 128 FMAs per iteration
    void* out = nullptr;
    CUDA_CHECK(cudaMalloc(&out, (size_t)threads * blocks * 8));
    cudaEvent_t a, b;
    CUDA_CHECK(cudaEventCreate(&a));
    CUDA_CHECK(cudaEventCreate(&b));
    float best = 1e30f;
    for (int rep = 0; rep < 4; ++rep) {
      CUDA_CHECK(cudaEventRecord(a));
      if (dtype == PASIF_F32) fma_peak_kernel<float><<<blocks, threads>>>((float*)out, iters, 0.999f, 1e-3f);
      else if (dtype == PASIF_F64) fma_peak_kernel<double><<<blocks, threads>>>((double*)out, iters, 0.999, 1e-3);
      else if (dtype == 2) dmma_peak_kernel<<<blocks, threads>>>((double*)out, iters * 2);
      else if (dtype == 3) tf32_mma_peak_kernel<<<blocks, threads>>>((float*)out, iters * 2);
      else if (dtype == 4) fma2_peak_kernel<<<blocks, threads>>>((float2*)out, iters / 2, 0.999f, 1e-3f);
      else throw std::runtime_error("pasif_fma_peak: dtype must be 0 (f64) | 1 (f32) | 2 (dmma) | 3 (tf32 mma.sync) | 4 (f32x2)");
      CUDA_CHECK(cudaEventRecord(b));
      CUDA_CHECK(cudaEventSynchronize(b));
      float ms = 0;
      CUDA_CHECK(cudaEventElapsedTime(&ms, a, b));
      if (rep > 0) best = std::min(best, ms);
    }
    CUDA_CHECK(cudaGetLastError());
    cudaEventDestroy(a); cudaEventDestroy(b); cudaFree(out);
    if (dtype == 2) *tflops = 2.0 * 256.0 * 4.0 * (double)(iters * 2) * (threads / 32) * blocks / (best * 1e-3) / 1e12;        // m8n8k4
    else if (dtype == 3) *tflops = 2.0 * 1024.0 * 4.0 * (double)(iters * 2) * (threads / 32) * blocks / (best * 1e-3) / 1e12;  // m16n8k8
    else *tflops = 2.0 * 128.0 * (double)iters * threads * blocks / (best * 1e-3) / 1e12;     // f32x2: iters / 2 x 256 FMAs
  } catch (const std::exception& ex) {
    g_create_error = ex.what();
    return 2;
  }
  return 0;
}

}  // extern "C"
