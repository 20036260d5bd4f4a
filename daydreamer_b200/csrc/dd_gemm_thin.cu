// Thin GEMMs: C[M,N] = A[M,K] op(B) (+bias) + beta*C with many rows but a tiny N (<= 32:
// action heads N = A or 2A, value / reward / cont outputs N = 1) or a tiny K (<= 32: the
// action part of img_in, K = A; input gradients of the N = 1 heads, K = 1).  Neither fits
// a tensor-core tile (the 128 x 64 tile would be > 80% padding), and the generic SIMT
// tile kernel spends ~27 us on each (r01 launch list: 79 launches, 7% of the a1 step).
// These are HBM/L2 streaming kernels: the small operand lives in shared memory, rows stream.
//   thin-N, N <= 4 : one warp per row, lanes split K (float4), warp reduction per column
//   thin-N, N <= 32: one warp per 4 rows, lane = column, A broadcast by shuffle
//   thin-K, K <= 32: one thread per 4 rows x 4 columns, B in shared memory
// Reference: nets.py:566-579 (Linear) for the shapes of DistLayer outputs (nets.py:447-492).
#include "dd_common.cuh"

namespace dd {

// Off by default: measured inside the captured a1 step (r01, profiles/r01_experiments.txt) the
// generic SIMT tile kernel is FASTER on these shapes than this first version (step 28.8 ms vs
// 29.2 ms with the thin kernels); the 27 us per launch that motivated them is the cold-cache,
// serialised ncu figure, not the in-graph cost.  Kept (and tested) behind dd_set_gemm_thin(1).
int g_thin_on = 2;

constexpr int THIN_SMEM = 96 * 1024;

// B as [K][NP] in shared memory (NP = padded N), from either operand major.
__device__ __forceinline__ void stage_b(float* Bs, const float* __restrict__ B, int ldb, int tb,
                                        int K, int N, int NP) {
  for (int e = threadIdx.x; e < K * NP; e += blockDim.x) {
    const int k = e / NP, n = e % NP;
    Bs[e] = n < N ? (tb ? B[(size_t)n * ldb + k] : B[(size_t)k * ldb + n]) : 0.f;
  }
}

template <int NN>     // N <= NN; NN in {1, 4} (level 1) or {8, 16} (level 2, experimental)
__global__ void __launch_bounds__(256) thin_n_small_kernel(
    const float* __restrict__ A, int lda, const float* __restrict__ B, int ldb, int tb,
    float* __restrict__ C, int ldc, int M, int N, int K, float beta,
    const float* __restrict__ bias) {
  pdl_prologue();
  extern __shared__ __align__(16) float thin_smem[];
  float* Bs = thin_smem;                                   // [N][K] (row n contiguous over k)
  if (NN > 4 && !tb) {
    // (level 2) B is [K, N] row-major: walk it in memory order so the global reads coalesce
    for (int e = threadIdx.x; e < K * N; e += blockDim.x) {
      const int k = e / N, n = e % N;
      Bs[n * K + k] = B[(size_t)k * ldb + n];
    }
  } else {
    for (int e = threadIdx.x; e < K * N; e += blockDim.x) {
      const int n = e / K, k = e % K;
      Bs[e] = tb ? B[(size_t)n * ldb + k] : B[(size_t)k * ldb + n];
    }
  }
  __syncthreads();
  const int lane = threadIdx.x & 31, wpb = blockDim.x >> 5;
  const bool vec = (lda & 3) == 0 && (K & 3) == 0 && (reinterpret_cast<uintptr_t>(A) & 15) == 0;
  for (int64_t m = blockIdx.x * (int64_t)wpb + (threadIdx.x >> 5); m < M;
       m += (int64_t)gridDim.x * wpb) {
    const float* a = A + (size_t)m * lda;
    float acc[NN];
#pragma unroll
    for (int n = 0; n < NN; ++n) acc[n] = 0.f;
    if (vec) {
      for (int k = lane * 4; k < K; k += 128) {
        const float4 av = *reinterpret_cast<const float4*>(a + k);
#pragma unroll
        for (int n = 0; n < NN; ++n) {
          if (n < N) {
            const float4 bv = *reinterpret_cast<const float4*>(Bs + n * K + k);
            acc[n] = fmaf(av.x, bv.x, fmaf(av.y, bv.y, fmaf(av.z, bv.z, fmaf(av.w, bv.w, acc[n]))));
          }
        }
      }
    } else {
      for (int k = lane; k < K; k += 32) {
        const float av = a[k];
#pragma unroll
        for (int n = 0; n < NN; ++n)
          if (n < N) acc[n] = fmaf(av, Bs[n * K + k], acc[n]);
      }
    }
#pragma unroll
    for (int n = 0; n < NN; ++n) acc[n] = warp_sum(acc[n]);
    if (lane == 0) {
#pragma unroll
      for (int n = 0; n < NN; ++n) {
        if (n < N) {
          float v = acc[n];
          if (bias != nullptr) v += bias[n];
          float* cp = C + (size_t)m * ldc + n;
          *cp = beta != 0.f ? v + beta * (*cp) : v;
        }
      }
    }
  }
}

__global__ void __launch_bounds__(256) thin_n_kernel(
    const float* __restrict__ A, int lda, const float* __restrict__ B, int ldb, int tb,
    float* __restrict__ C, int ldc, int M, int N, int K, int KC, float beta,
    const float* __restrict__ bias) {
  pdl_prologue();
  extern __shared__ __align__(16) float thin_smem[];
  float* Bs = thin_smem;                                   // [KC][32]
  const int lane = threadIdx.x & 31, wpb = blockDim.x >> 5;
  const int64_t groups = ((int64_t)M + 3) / 4;
  const int64_t g0 = blockIdx.x * (int64_t)wpb + (threadIdx.x >> 5);
  const int64_t gstep = (int64_t)gridDim.x * wpb;
  for (int k0 = 0; k0 < K; k0 += KC) {
    const int kc = min(KC, K - k0);
    __syncthreads();
    stage_b(Bs, tb ? B + k0 : B + (size_t)k0 * ldb, ldb, tb, kc, N, 32);
    __syncthreads();
    for (int64_t g = g0; g < groups; g += gstep) {
      const int64_t m0 = g * 4;
      float acc[4] = {0.f, 0.f, 0.f, 0.f};
      for (int kb = 0; kb < kc; kb += 32) {
        float ar[4];
#pragma unroll
        for (int r = 0; r < 4; ++r)
          ar[r] = (m0 + r < M && kb + lane < kc) ? A[(size_t)(m0 + r) * lda + k0 + kb + lane] : 0.f;
        const int kn = min(32, kc - kb);
        for (int kk = 0; kk < kn; ++kk) {
          const float bv = Bs[(kb + kk) * 32 + lane];
#pragma unroll
          for (int r = 0; r < 4; ++r)
            acc[r] = fmaf(__shfl_sync(0xffffffffu, ar[r], kk), bv, acc[r]);
        }
      }
      if (lane < N) {
#pragma unroll
        for (int r = 0; r < 4; ++r) {
          if (m0 + r < M) {
            float* cp = C + (size_t)(m0 + r) * ldc + lane;
            float v = acc[r];
            if (k0 == 0) {
              if (bias != nullptr) v += bias[lane];
              *cp = beta != 0.f ? v + beta * (*cp) : v;
            } else {
              *cp += v;
            }
          }
        }
      }
    }
  }
}

__global__ void __launch_bounds__(256) thin_k_kernel(
    const float* __restrict__ A, int lda, const float* __restrict__ B, int ldb, int tb,
    float* __restrict__ C, int ldc, int M, int N, int K, float beta,
    const float* __restrict__ bias) {
  pdl_prologue();
  extern __shared__ __align__(16) float thin_smem[];
  const int NP = (N + 3) / 4 * 4;
  float* Bs = thin_smem;                                   // [K][NP]
  stage_b(Bs, B, ldb, tb, K, N, NP);
  __syncthreads();
  const int n4 = NP / 4;
  const int64_t groups = ((int64_t)M + 3) / 4;
  const int64_t total = groups * n4;
  for (int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; e < total;
       e += (int64_t)gridDim.x * blockDim.x) {
    const int c = (int)(e % n4) * 4;
    const int64_t m0 = (e / n4) * 4;
    float acc[4][4];
#pragma unroll
    for (int r = 0; r < 4; ++r)
#pragma unroll
      for (int j = 0; j < 4; ++j) acc[r][j] = 0.f;
    for (int k = 0; k < K; ++k) {
      const float4 bv = *reinterpret_cast<const float4*>(Bs + k * NP + c);
#pragma unroll
      for (int r = 0; r < 4; ++r) {
        const float av = m0 + r < M ? A[(size_t)(m0 + r) * lda + k] : 0.f;
        acc[r][0] = fmaf(av, bv.x, acc[r][0]);
        acc[r][1] = fmaf(av, bv.y, acc[r][1]);
        acc[r][2] = fmaf(av, bv.z, acc[r][2]);
        acc[r][3] = fmaf(av, bv.w, acc[r][3]);
      }
    }
#pragma unroll
    for (int r = 0; r < 4; ++r) {
      if (m0 + r >= M) continue;
      float* cp = C + (size_t)(m0 + r) * ldc + c;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        if (c + j < N) {
          float v = acc[r][j];
          if (bias != nullptr) v += bias[c + j];
          cp[j] = beta != 0.f ? v + beta * cp[j] : v;
        }
      }
    }
  }
}

// 1 = handled, 0 = not eligible
int gemm_thin_try(const float* A, int lda, int ta, const float* B, int ldb, int tb, float* C,
                  int ldc, int M, int N, int K, float beta, const float* bias, cudaStream_t st) {
  if (!g_thin_on || ta || M <= 32 || K < 1) return 0;
  static bool once = false;
  if (!once) {
    cudaFuncSetAttribute(thin_n_small_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, THIN_SMEM);
    cudaFuncSetAttribute(thin_n_small_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, THIN_SMEM);
    cudaFuncSetAttribute(thin_n_small_kernel<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, THIN_SMEM);
    cudaFuncSetAttribute(thin_n_small_kernel<16>, cudaFuncAttributeMaxDynamicSharedMemorySize, THIN_SMEM);
    cudaFuncSetAttribute(thin_n_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, THIN_SMEM);
    cudaFuncSetAttribute(thin_k_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, THIN_SMEM);
    once = true;
  }
  // level 2 (experimental, written after the r01 measurement, not yet run): the row-per-warp kernel
  // also takes N <= 16 (the action heads), where the 4-rows-per-warp shuffle kernel measured slow
  const int nsmall = g_thin_on >= 2 ? 16 : 4;
  if (N <= nsmall && (size_t)K * N * 4 <= THIN_SMEM && K >= 16) {
    const int grid = grid_for(M, 8, 148 * 8);
    const size_t smem = (size_t)K * N * 4;
#define DD_THIN_S(NN) launch_k(thin_n_small_kernel<NN>, grid, 256, smem, st, A, lda, B, ldb, tb, C, ldc, M, N, K, beta, bias)
    if (N == 1) DD_THIN_S(1);
    else if (N <= 4) DD_THIN_S(4);
    else if (N <= 8) DD_THIN_S(8);
    else DD_THIN_S(16);
#undef DD_THIN_S
    return launched("dd_gemm_thin_n") == 0 ? 1 : -1;
  }
  if (K <= 32 && (size_t)K * ((N + 3) / 4 * 4) * 4 <= THIN_SMEM) {
    const int NP = (N + 3) / 4 * 4;
    const int64_t total = (((int64_t)M + 3) / 4) * (NP / 4);
    launch_k(thin_k_kernel, grid_for(total, 256, 148 * 8), 256, (size_t)K * NP * 4, st, A, lda, B, ldb, tb,
             C, ldc, M, N, K, beta, bias);
    return launched("dd_gemm_thin_k") == 0 ? 1 : -1;
  }
  if (N <= 32 && K >= 32) {
    const int KC = min((K + 31) / 32 * 32, THIN_SMEM / (32 * 4) / 32 * 32);
    const int64_t groups = ((int64_t)M + 3) / 4;
    launch_k(thin_n_kernel, grid_for(groups, 8, 148 * 4), 256, (size_t)KC * 32 * 4, st, A, lda, B, ldb, tb,
             C, ldc, M, N, K, KC, beta, bias);
    return launched("dd_gemm_thin_n") == 0 ? 1 : -1;
  }
  return 0;
}

}  // namespace dd

extern "C" int dd_set_gemm_thin(int on) {
  const int prev = dd::g_thin_on;
  dd::g_thin_on = on < 0 ? 0 : (on > 2 ? 2 : on);       // 0 off, 1 r01 kernels, 2 + row-per-warp up to N = 16
  return prev;
}
