// Shared device/host helpers for the dd_* kernel library (sm_100a).
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include <atomic>

#include "../../include/dd_kernels.h"

#define DD_LN_EPS 1e-3f

namespace dd {

extern thread_local char g_err[256];
extern std::atomic<int64_t> g_launches;

inline int fail(const char* what, const char* detail) {
  snprintf(g_err, sizeof(g_err), "%s: %s", what, detail);
  return -1;
}

// Called after every launch: counts it and surfaces launch-configuration errors.
inline int launched(const char* what, int n = 1) {
  g_launches.fetch_add(n, std::memory_order_relaxed);
  cudaError_t e = cudaPeekAtLastError();
  if (e != cudaSuccess) {
    cudaGetLastError();
    return fail(what, cudaGetErrorString(e));
  }
  return 0;
}

inline cudaStream_t S(dd_stream_t s) { return reinterpret_cast<cudaStream_t>(s); }

// Programmatic dependent launch (dd_set_pdl).  Every kernel starts with
// pdl_prologue(): it lets the next kernel of the stream be scheduled as soon as all
// CTAs of this one are resident, then blocks until the previous kernel has
// completed and its memory is visible -- so the data dependence between
// consecutive launches is unchanged and only launch / CTA-scheduling latency
// overlaps.  Without the launch attribute both instructions are no-ops.
extern int g_pdl;

__device__ __forceinline__ void pdl_prologue() {
  asm volatile("griddepcontrol.launch_dependents;");
  asm volatile("griddepcontrol.wait;" ::: "memory");
}

template <typename... KArgs, typename... Args>
inline void launch_k(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st,
                     Args&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  at[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = at;
  cfg.numAttrs = g_pdl ? 1 : 0;
  cudaLaunchKernelEx(&cfg, kern, static_cast<KArgs>(args)...);
}

inline int grid_for(int64_t n, int block, int cap = 148 * 16) {
  int64_t g = (n + block - 1) / block;
  if (g < 1) g = 1;
  if (g > cap) g = cap;
  return static_cast<int>(g);
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

__device__ __forceinline__ double warp_sum_d(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// Block-wide sum of one float; every thread gets the result.  `red` must hold
// >= 33 floats.  Safe to call repeatedly (leading __syncthreads).
__device__ __forceinline__ float block_sum(float v, float* red) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int nwarp = (blockDim.x + 31) >> 5;
  v = warp_sum(v);
  __syncthreads();
  if (lane == 0) red[warp] = v;
  __syncthreads();
  if (warp == 0) {
    float t = lane < nwarp ? red[lane] : 0.f;
    t = warp_sum(t);
    if (lane == 0) red[32] = t;
  }
  __syncthreads();
  return red[32];
}

__device__ __forceinline__ float sigmoidf_(float x) { return 1.f / (1.f + expf(-x)); }

__device__ __forceinline__ float symlogf_(float x) {
  return copysignf(log1pf(fabsf(x)), x);
}

__device__ __forceinline__ float symexpf_(float x) {
  return copysignf(expm1f(fabsf(x)), x);
}

__device__ __forceinline__ double atomic_min_d(double* addr, double val) {
  unsigned long long* a = reinterpret_cast<unsigned long long*>(addr);
  unsigned long long old = *a, assumed;
  do {
    assumed = old;
    if (__longlong_as_double(assumed) <= val) break;
    old = atomicCAS(a, assumed, __double_as_longlong(val));
  } while (assumed != old);
  return __longlong_as_double(old);
}

__device__ __forceinline__ double atomic_max_d(double* addr, double val) {
  unsigned long long* a = reinterpret_cast<unsigned long long*>(addr);
  unsigned long long old = *a, assumed;
  do {
    assumed = old;
    if (__longlong_as_double(assumed) >= val) break;
    old = atomicCAS(a, assumed, __double_as_longlong(val));
  } while (assumed != old);
  return __longlong_as_double(old);
}

}  // namespace dd
