/* pasif_b200.h -- C ABI of the B200-native RK4 forward / adjoint engine.
 *
 * Drop-in boundary: these entry points are what the reference's pybind11 class `PASIfgpu.__GpuDriver`
 * (reference src/PASIfgpu.cpp:22-105, declared in src/__GpuDriver.cuh:18-110) binds; each function cites the
 * reference method it replaces.  Plain pointers and sizes only.  Every function returns 0 on success and a
 * non-zero status on failure; pasif_last_error() then describes the failure (the reference prints and calls
 * std::exit(0) instead, src/helpers.cuh:43-51).  Inputs are always double precision, as the reference's Python
 * front end passes them; the engine rounds them to its compute type (`reel` in the reference).
 *
 * Index conventions are the reference's: CSR K' with sorted column indices; COO tensors with flat indices
 * [out, in1, in2(, in3(, in4))] per non-zero, where out/in1 index the differentiated state and in2.. index the
 * forward state (src/kernels.cu:52,56,61); excitations row-major [E][L]; results row-major [e][i].
 */
#ifndef PASIF_B200_H
#define PASIF_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct pasif_engine pasif_engine;

enum { PASIF_F64 = 0, PASIF_F32 = 1 };

/* __GpuDriver::__GpuDriver (src/__GpuDriver.cu:25-157) + setCUDA (:595-630) + _loadExcitationsSet (:343-381).
 * dtype: PASIF_F64 / PASIF_F32 (compile-time `#define reel` in the reference, src/helpers.cuh:29).
 * device: CUDA device ordinal (GPUId_); -1 creates a host-only engine that can analyse systems but not solve. */
int pasif_create(pasif_engine** out, const double* excitations, uint32_t n_excitations, uint32_t length,
                 uint32_t sample_rate, uint32_t numsteps, int dtype, int device);
/* __GpuDriver::~__GpuDriver (src/__GpuDriver.cu:159-193) */
void pasif_destroy(pasif_engine* e);
/* error text of the last failing call on this engine (e == NULL: of the last failing pasif_create) */
const char* pasif_last_error(const pasif_engine* e);

/* __GpuDriver::_loadExcitationsSet (src/__GpuDriver.cu:343-381) */
int pasif_load_excitations(pasif_engine* e, const double* excitations, uint32_t n_excitations, uint32_t length,
                           uint32_t sample_rate);

/* forward system: _setFwdK/_setFwdGamma/_setFwdLambda/_setFwdPsi/_setFwdForcePattern/_setFwdInitialConditions
 * (src/__GpuDriver.cu:197-252) and _allocateSystemOnDevice (:256-265) */
int pasif_set_fwd_k(pasif_engine* e, uint32_t n_rows, uint32_t n_cols, const double* values, const uint32_t* indices,
                    const uint32_t* indptr, uint32_t nnz);
int pasif_set_fwd_gamma(pasif_engine* e, const uint32_t dims[3], const double* values, const uint32_t* indices, uint32_t nnz);
int pasif_set_fwd_lambda(pasif_engine* e, const uint32_t dims[4], const double* values, const uint32_t* indices, uint32_t nnz);
int pasif_set_fwd_psi(pasif_engine* e, const uint32_t dims[5], const double* values, const uint32_t* indices, uint32_t nnz);
int pasif_set_fwd_force(pasif_engine* e, const double* dense, uint32_t n);
int pasif_set_fwd_ic(pasif_engine* e, const double* ic, uint32_t n);
int pasif_alloc_system(pasif_engine* e);

/* adjoint system: _setBwd* (src/__GpuDriver.cu:272-327) and _allocateAdjointOnDevice (:329-338);
 * tensor dims are [N_adj, N_adj, N_fwd(, N_fwd(, N_fwd))] */
int pasif_set_bwd_k(pasif_engine* e, uint32_t n_rows, uint32_t n_cols, const double* values, const uint32_t* indices,
                    const uint32_t* indptr, uint32_t nnz);
int pasif_set_bwd_gamma(pasif_engine* e, const uint32_t dims[3], const double* values, const uint32_t* indices, uint32_t nnz);
int pasif_set_bwd_lambda(pasif_engine* e, const uint32_t dims[4], const double* values, const uint32_t* indices, uint32_t nnz);
int pasif_set_bwd_psi(pasif_engine* e, const uint32_t dims[5], const double* values, const uint32_t* indices, uint32_t nnz);
int pasif_set_bwd_force(pasif_engine* e, const double* dense, uint32_t n);   /* stored, unused: the adjoint ODE is unforced (:756-780) */
int pasif_set_bwd_ic(pasif_engine* e, const double* ic, uint32_t n);
int pasif_alloc_adjoint(pasif_engine* e);

/* _setInterpolationMatrix (src/__GpuDriver.cu:383-398): table = 2 x window row-major, window a power of two */
int pasif_set_interp(pasif_engine* e, const double* table, uint32_t table_len, uint32_t window);
/* additive: the 2 x W table of LINEAR interpolation between consecutive samples, generated on the device (what a caller of
 * the reference builds on the host and passes to _setInterpolationMatrix; layout read by interpolateForces_fwd,
 * src/kernels.cu:114-124: IM[j] = j / W weighs exc[o + 1], IM[W + j] = 1 - j / W weighs exc[o]).  W must be a power of two. */
int pasif_set_interp_linear(pasif_engine* e, uint32_t window);
/* _setModulationBuffer (src/__GpuDriver.cu:400-410): accepted and stored; no reference kernel reads it */
int pasif_set_modulation(pasif_engine* e, const double* buffer, uint32_t n);

/* engine options (no reference counterpart): "compat" = "reference" | "clean" (adjoint schedule semantics),
 * "jit" = "on" | "off" (allow NVRTC specialisation for patterns without a prebuilt kernel), "block" = threads per CTA,
 * "lanes" = "on" | "off" | widest thread group (8..128) of the lane-block split of large components,
 * "large_threshold", "gemm", "tcgen05", "overlap", "profile", "tail_budget_gb", "numsteps" */
int pasif_set_option(pasif_engine* e, const char* key, const char* value);

/* _getAmplitudes (src/__GpuDriver.cu:416-461): out[e * N + i], n_out = N * E.  out == NULL runs the solve and keeps
 * the result on the device (throughput measurement with resident inputs). */
int pasif_get_amplitudes(pasif_engine* e, double* out, uint64_t n_out);
/* _getTrajectory (src/__GpuDriver.cu:463-500): out[(t * E + e) * N + i]; n_out = pasif_trajectory_size() */
uint64_t pasif_trajectory_size(const pasif_engine* e, uint32_t save_steps);
int pasif_get_trajectory(pasif_engine* e, uint32_t save_steps, double* out, uint64_t n_out);
/* _getGradient (src/__GpuDriver.cu:502-587): out[e * N_adj + j], n_out = N_adj * E; `save` stops after that chunk */
int pasif_get_gradient(pasif_engine* e, uint32_t save, double* out, uint64_t n_out);
/* additive API: adjoint state summed over the excitation axis on the device (warp-shuffle reduction), out[j], j < N_adj.
 * The _device variant writes N_adj values of the engine's compute type to a device pointer (for an NCCL all-reduce). */
int pasif_get_gradient_reduced(pasif_engine* e, uint32_t save, double* out, uint64_t n_out);
int pasif_get_gradient_reduced_device(pasif_engine* e, uint32_t save, void* device_out, uint64_t n_out);
/* the same sum over excitations of the adjoint state the LAST pasif_get_gradient call left on the device (no recompute):
 * lets a caller fetch the reference's un-reduced [e][j] result and feed the all-reduce from one evaluation */
int pasif_reduce_last_gradient_device(pasif_engine* e, void* device_out, uint64_t n_out);

/* introspection: "n_fwd", "n_adj", "n_excitations", "length", "numsteps", "fwd_classes", "fwd_instances",
 * "joint_classes", "joint_instances", "last_device_ms" (CUDA-event time of the last solve on the engine's stream),
 * "last_launches" (kernels launched by the last solve), "jit_compiles", "aot_hits", "dtype" */
int pasif_query(const pasif_engine* e, const char* key, double* value);
/* C++ source of the compile-time pattern traits of the currently loaded systems (used by tools/gen_aot_patterns.py
 * to pre-build kernels); writes at most cap bytes incl. NUL, *needed = bytes required */
int pasif_emit_patterns(pasif_engine* e, char* buf, uint64_t cap, uint64_t* needed);
/* NVRTC + ptxas -v logs of the run-time specialised kernels of the loaded classes, compiled for `arch` ("sm_100a");
 * no device needed (debugging aid: register use of the lane-block kernels) */
int pasif_debug_jit_log(pasif_engine* e, const char* arch, char* buf, uint64_t cap, uint64_t* needed);
/* device pipe microbenchmarks (register-resident, no memory traffic): achieved TFLOP/s.  dtype PASIF_F64 / PASIF_F32: FMA pipe;
 * 2: mma.sync m8n8k4 f64 (DMMA); 3: mma.sync m16n8k8 tf32 (fp32 accumulate); 4: FMA pipe through packed fp32 pairs (FFMA2).
 * The FMA loops carry 128 FMAs per iteration (>= 97 % of the issued instructions) */
int pasif_fma_peak(int device, int dtype, double* tflops);
/* isolated check of the dense-operator path's GEMM kernels (no reference counterpart: replaces the cusparseSpMM of
 * src/__GpuDriver.cu:795-807 for a dense shared K'): y[p][m] = sum_k kd[m][k] * x[p][k], kd [n][n], x / y [n_traj][n], row-major doubles.
 * mode 0: mma.sync kernel of the dtype (fp64 DMMA, fp32 3xTF32); mode 1: tcgen05 3xTF32 with pre-split operands (fp32 only);
 * mode 2: tcgen05 3xTF32 that splits the fp32 operand in shared memory (fp32 only); mode 4: the same with 256-row tiles;
 * modes 5 / 6: modes 2 / 4 with 16-float k-blocks (64-byte swizzle, twice the pipeline stages) */
int pasif_debug_gemm(int device, int dtype, int mode, uint32_t n, uint64_t n_traj, const double* kd, const double* x, double* y);

#ifdef __cplusplus
}
#endif
#endif
