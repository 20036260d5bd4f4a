// Microbenchmark: the layer machine's MMA k-loop in isolation.  One elected thread per CTA issues,
// per k-block, 4 k-steps x { A_hi x [B_hi;B_lo] (N = 128) , A_lo x B_hi (N = 64) } on kind::tf32 and
// then commits to an mbarrier -- local or multicast to the 8 CTAs of a cluster -- exactly like
// mma_issuer in csrc/dd_machine.cu, but with no producers (operands are whatever is in shared memory).
// Question: why does the machine's loop take ~1100 cycles per k-block when 8 MMAs cost 8 x 75?
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o mma_loop mma_loop.cu && ./mma_loop
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
__device__ __forceinline__ uint32_t make_idesc(int n) {
  return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
}
__device__ __forceinline__ void mma(uint32_t d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t acc) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
               "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
               ::"r"(d), "l"(da), "l"(db), "r"(idesc), "r"(acc) : "memory");
}

enum { M_FAST = 64,        // lean issue loop: descriptors precomputed per stage, k-steps unrolled, no modulo
       M_UNIFORM = 128,    // second product is A_lo x [B_hi;B_lo] (N = 128, same shape as the first)
       M_COMMIT = 1,      // commit after every k-block (else only once at the end)
       M_MC = 2,          // ... multicast to all 8 CTAs of the cluster (barrier count 8)
       M_SEPARATE = 4,    // the N = 64 product accumulates into its own TMEM columns
       M_ONLY1 = 8,       // only the N = 128 products
       M_WAIT = 16,       // wait for k-block (kb - 3)'s commit before issuing k-block kb (ring of 3)
       M_N256 = 32 };     // one N = 256 product per k-step instead (peak-rate shape)

constexpr int STAGE = 49152;

__global__ void __launch_bounds__(128, 1) loop_kernel(int mode, int kblocks, long long* out) {
  extern __shared__ __align__(1024) char smem[];
  __shared__ __align__(8) uint64_t bars[4];
  __shared__ uint32_t tmem_slot;
  const int tid = threadIdx.x, warp = tid >> 5;
  uint32_t rank, nctas;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(rank));
  asm volatile("mov.u32 %0, %%cluster_nctarank;" : "=r"(nctas));
  const bool mc = (mode & M_MC) != 0;
  if (tid == 0) {
    for (int i = 0; i < 4; ++i)
      asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(&bars[i])), "r"(mc ? nctas : 1u) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;"
                 ::"r"(smem_u32(&tmem_slot)), "r"(512u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = tmem_slot;
  long long t0 = 0, t1 = 0, t2 = 0;
  if (tid == 0) {
    const uint32_t base = smem_u32(smem);
    const uint32_t i64 = make_idesc(64), i128 = make_idesc(128), i256 = make_idesc(256);
    const uint16_t mask = (uint16_t)((1u << nctas) - 1u);
    t0 = clock64();
    for (int kb = 0; kb < kblocks; ++kb) {
      const uint32_t s = kb % 3;
      if ((mode & M_WAIT) && kb >= 3) {
        const uint32_t ph = ((kb / 3) - 1) & 1;
        uint32_t done = 0;
        while (!done)
          asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
                       "selp.b32 %0, 1, 0, p;\n\t}" : "=r"(done) : "r"(smem_u32(&bars[s])), "r"(ph) : "memory");
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      }
      const uint32_t a_hi = base + s * STAGE, a_lo = a_hi + 16384, b = (mode & M_N256) ? a_hi : a_hi + 32768;
      for (int ks = 0; ks < 4; ++ks) {
        const uint64_t dah = make_desc(a_hi + ks * 32), dal = make_desc(a_lo + ks * 32), db = make_desc(b + ks * 32);
        const uint32_t acc = (kb | ks) != 0;
        if (mode & M_N256) { mma(tmem, dah, db, i256, acc); continue; }
        mma(tmem, dah, db, i128, acc);
        if (!(mode & M_ONLY1)) mma(tmem + ((mode & M_SEPARATE) ? 128u : 64u), dal, db, i64, (mode & M_SEPARATE) ? acc : 1u);
      }
      if (mode & M_COMMIT) {
        if (mc)
          asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
                       ::"r"(smem_u32(&bars[s])), "h"(mask) : "memory");
        else
          asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];"
                       ::"r"(smem_u32(&bars[s])) : "memory");
      }
    }
    t1 = clock64();
    // drain: a final local commit on a barrier nobody else touches
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bars[3])) : "memory");
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];"
                 ::"r"(smem_u32(&bars[3])) : "memory");
    uint32_t done = 0;
    while (!done)
      asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
                   "selp.b32 %0, 1, 0, p;\n\t}" : "=r"(done) : "r"(smem_u32(&bars[3])), "r"(0u) : "memory");
    t2 = clock64();
    if (blockIdx.x == 0) { out[0] = t1 - t0; out[1] = t2 - t0; }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
  if (warp == 0)
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512u) : "memory");
}

// The same loop written the way the issuing thread should run it: everything that does not change
// per k-block is hoisted, the stage index is a counter, the four k-steps are unrolled.
template <int MODE>
__global__ void __launch_bounds__(128, 1) fast_kernel(int kblocks, long long* out) {
  extern __shared__ __align__(1024) char smem[];
  __shared__ __align__(8) uint64_t bars[4];
  __shared__ uint32_t tmem_slot;
  const int tid = threadIdx.x, warp = tid >> 5;
  uint32_t nctas;
  asm volatile("mov.u32 %0, %%cluster_nctarank;" : "=r"(nctas));
  constexpr bool mc = (MODE & M_MC) != 0;
  if (tid == 0) {
    for (int i = 0; i < 4; ++i)
      asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(&bars[i])), "r"(mc ? nctas : 1u) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;"
                 ::"r"(smem_u32(&tmem_slot)), "r"(512u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = tmem_slot;
  long long t0 = 0, t1 = 0, t2 = 0;
  if (tid == 0) {
    const uint32_t base = smem_u32(smem);
    const uint32_t i64 = make_idesc(64), i128 = make_idesc(128);
    const uint16_t mask = (uint16_t)((1u << nctas) - 1u);
    uint64_t dah[3], dal[3], db[3];
    uint32_t bar[3];
#pragma unroll
    for (int s = 0; s < 3; ++s) {
      dah[s] = make_desc(base + s * STAGE); dal[s] = make_desc(base + s * STAGE + 16384);
      db[s] = make_desc(base + s * STAGE + 32768); bar[s] = smem_u32(&bars[s]);
    }
    uint32_t ph = 0;
    t0 = clock64();
    for (int kb = 0; kb < kblocks; kb += 3) {
#pragma unroll
      for (int s = 0; s < 3; ++s) {
        if (kb + s >= kblocks) break;
        if ((MODE & M_WAIT) && kb >= 3) {
          uint32_t done = 0;
          while (!done)
            asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
                         "selp.b32 %0, 1, 0, p;\n\t}" : "=r"(done) : "r"(bar[s]), "r"(ph ^ 1u) : "memory");
          asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        }
#pragma unroll
        for (int ks = 0; ks < 4; ++ks) {
          const uint32_t acc = (kb | s | ks) != 0;
          // +32 bytes per k-step = +2 in the descriptor's 16-byte address units
          mma(tmem, dah[s] + 2 * ks, db[s] + 2 * ks, i128, acc);
          if (MODE & M_UNIFORM) mma(tmem, dal[s] + 2 * ks, db[s] + 2 * ks, i128, 1u);
          else if (!(MODE & M_ONLY1)) mma(tmem + 64u, dal[s] + 2 * ks, db[s] + 2 * ks, i64, 1u);
        }
        if (MODE & M_COMMIT) {
          if (mc)
            asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
                         ::"r"(bar[s]), "h"(mask) : "memory");
          else
            asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];"
                         ::"r"(bar[s]) : "memory");
        }
      }
      ph ^= 1u;
    }
    t1 = clock64();
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bars[3])) : "memory");
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];"
                 ::"r"(smem_u32(&bars[3])) : "memory");
    uint32_t done = 0;
    while (!done)
      asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
                   "selp.b32 %0, 1, 0, p;\n\t}" : "=r"(done) : "r"(smem_u32(&bars[3])), "r"(0u) : "memory");
    t2 = clock64();
    if (blockIdx.x == 0) { out[0] = t1 - t0; out[1] = t2 - t0; }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
  if (warp == 0)
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512u) : "memory");
}

template <int MODE>
static void launch_fast(int cluster, int grid, int kblocks, long long* out) {
  cudaFuncSetAttribute(fast_kernel<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, 3 * STAGE);
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(grid); cfg.blockDim = dim3(128); cfg.dynamicSmemBytes = 3 * STAGE;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeClusterDimension;
  at[0].val.clusterDim.x = cluster; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
  cfg.attrs = at; cfg.numAttrs = 1;
  cudaLaunchKernelEx(&cfg, fast_kernel<MODE>, kblocks, out);
}

template <int MODE>
static void report_fast(int cluster, int grid, const char* what, long long* out) {
  const int kblocks = 255;
  long long h[2];
  for (int w = 0; w < 2; ++w) { launch_fast<MODE>(cluster, grid, kblocks, out); cudaDeviceSynchronize(); }
  cudaMemcpy(h, out, 16, cudaMemcpyDeviceToHost);
  printf("grid %3d %-62s issue %7.1f cyc/k-block, complete %7.1f cyc/k-block (%s)\n", grid, what,
         (double)h[0] / kblocks, (double)h[1] / kblocks, cudaGetErrorString(cudaGetLastError()));
}

static void launch(int cluster, int grid, int mode, int kblocks, long long* out) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(grid); cfg.blockDim = dim3(128); cfg.dynamicSmemBytes = 3 * STAGE;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeClusterDimension;
  at[0].val.clusterDim.x = cluster; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
  cfg.attrs = at; cfg.numAttrs = 1;
  cudaLaunchKernelEx(&cfg, loop_kernel, mode, kblocks, out);
}

int main() {
  setvbuf(stdout, NULL, _IONBF, 0);
  long long* out;
  cudaMalloc(&out, 16);
  cudaFuncSetAttribute(loop_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 3 * STAGE);
  const int kblocks = 256;
  struct { int cluster, mode; const char* what; } cases[] = {
      {1, 0, "8 MMAs/k-block, no commits"},
      {1, M_COMMIT, "8 MMAs + local commit"},
      {1, M_COMMIT | M_WAIT, "8 MMAs + local commit, ring of 3 waits"},
      {1, M_COMMIT | M_SEPARATE, "8 MMAs (separate cross columns) + local commit"},
      {1, M_COMMIT | M_ONLY1, "4 MMAs N=128 + local commit"},
      {1, M_COMMIT | M_ONLY1 | M_WAIT, "4 MMAs N=128 + local commit, ring waits"},
      {1, M_COMMIT | M_N256, "4 MMAs N=256 + local commit"},
  };
  for (int grid : {8, 56})
    for (auto& c : cases) {
      long long h[2];
      for (int w = 0; w < 2; ++w) { launch(c.cluster, grid, c.mode, kblocks, out); cudaDeviceSynchronize(); }
      cudaMemcpy(h, out, 16, cudaMemcpyDeviceToHost);
      printf("grid %3d %-62s issue %7.1f cyc/k-block, complete %7.1f cyc/k-block (%s)\n", grid, c.what,
             (double)h[0] / kblocks, (double)h[1] / kblocks, cudaGetErrorString(cudaGetLastError()));
    }
  for (int grid : {8, 56}) {
    report_fast<M_FAST>(1, grid, "lean: 8 MMAs, no commits", out);
    report_fast<M_FAST | M_COMMIT>(1, grid, "lean: 8 MMAs + local commit", out);
    report_fast<M_FAST | M_COMMIT | M_WAIT>(1, grid, "lean: 8 MMAs + local commit, ring waits", out);
    report_fast<M_FAST | M_COMMIT | M_WAIT | M_ONLY1>(1, grid, "lean: 4 MMAs N=128 + local commit, ring waits", out);
    report_fast<M_FAST | M_COMMIT | M_WAIT | M_UNIFORM>(1, grid, "lean: 8 MMAs all N=128 (4 products) + local commit, ring waits", out);
  }
  return 0;
}
