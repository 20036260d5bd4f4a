// Microbenchmark: cycles per tcgen05.mma (cta_group::1, M = 128) for kind::tf32 (K = 8) and
// kind::f16 / bf16 (K = 16), N in {64, 128, 256}, operands = K-major SWIZZLE_128B tiles in shared
// memory (contents irrelevant).  One CTA per SM, `reps` back-to-back MMAs, one commit at the end.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o mma_rate mma_rate.cu && ./mma_rate
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ uint64_t make_desc(uint32_t addr) {
  return (uint64_t)((addr & 0x3FFFFu) >> 4) | (1ull << 16) | ((uint64_t)(1024 >> 4) << 32) |
         (1ull << 46) | (2ull << 61);
}
// kind 0: tf32 (a,b format 2), kind 1: bf16 (format 1), c = f32
__device__ __forceinline__ uint32_t make_idesc(int kind, int n) {
  const uint32_t fmt = kind == 0 ? 2u : 1u;
  return (1u << 4) | (fmt << 7) | (fmt << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
}

template <int KIND>
__global__ void __launch_bounds__(128, 1) rate_kernel(int n, int reps, int same, long long* out) {
  extern __shared__ __align__(1024) char smem[];
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_slot;
  const int tid = threadIdx.x, warp = tid >> 5;
  if (tid == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;"
                 ::"r"(smem_u32(&tmem_slot)), "r"(512u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = tmem_slot;
  long long t0 = 0, t1 = 0, t2 = 0;
  if (tid == 0) {
    const uint32_t a0 = smem_u32(smem), b0 = a0 + 65536;
    const uint32_t idesc = make_idesc(KIND, n);
    t0 = clock64();
    for (int i = 0; i < reps; ++i) {
      // walk 4 k-steps x 4 stages of operand tiles like a real pipeline (same = 1: one tile only)
      const uint32_t off = same ? 0u : (uint32_t)(((i >> 2) & 3) * 16384 + (i & 3) * 32);
      const uint64_t da = make_desc(a0 + off), db = make_desc(b0 + (same ? 0u : (uint32_t)(((i >> 2) & 3) * 32768 + (i & 3) * 32)));
      const uint32_t accum = i > 0;
      if (KIND == 0)
        asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                     "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
                     ::"r"(tmem), "l"(da), "l"(db), "r"(idesc), "r"(accum) : "memory");
      else
        asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                     "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
                     ::"r"(tmem), "l"(da), "l"(db), "r"(idesc), "r"(accum) : "memory");
    }
    t1 = clock64();
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];"
                 ::"r"(smem_u32(&bar)) : "memory");
    uint32_t done = 0;
    while (!done)
      asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
                   "selp.b32 %0, 1, 0, p;\n\t}" : "=r"(done) : "r"(smem_u32(&bar)), "r"(0u) : "memory");
    t2 = clock64();
    if (blockIdx.x == 0) { out[0] = t1 - t0; out[1] = t2 - t0; }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0)
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512u) : "memory");
}

int main() {
  long long* out;
  cudaMalloc(&out, 16);
  const int smem = 65536 + 131072;
  cudaFuncSetAttribute(rate_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  cudaFuncSetAttribute(rate_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  const int reps = 512;
  for (int grid : {1, 148})
    for (int kind = 0; kind < 2; ++kind)
      for (int n : {64, 128, 256})
        for (int same = 0; same < 2; ++same) {
          long long h[2];
          for (int w = 0; w < 2; ++w) {
            if (kind == 0) rate_kernel<0><<<grid, 128, smem>>>(n, reps, same, out);
            else rate_kernel<1><<<grid, 128, smem>>>(n, reps, same, out);
            cudaDeviceSynchronize();
          }
          cudaMemcpy(h, out, 16, cudaMemcpyDeviceToHost);
          const int k = kind == 0 ? 8 : 16;
          printf("grid %3d %s M=128 N=%3d K=%2d %s: issue %.1f cyc/mma, complete %.1f cyc/mma  -> %.0f MAC/clk/SM  (%s)\n",
                 grid, kind == 0 ? "tf32" : "bf16", n, k, same ? "one tile " : "walk tiles", (double)h[0] / reps,
                 (double)h[1] / reps, 128.0 * n * k * reps / (double)h[1], cudaGetErrorString(cudaGetLastError()));
        }
  return 0;
}
