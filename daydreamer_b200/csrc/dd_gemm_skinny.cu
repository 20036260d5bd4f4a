// Skinny GEMMs of the observe scan: C[M,N] = A[M,K] op(B) (+bias) + beta*C with
// M = sequences per GPU (<= 32).  These are latency / L2-bandwidth bound weight
// streams, not FLOP bound.  The activation block A (M x K, <= 200 KB) is staged
// in shared memory ONCE, then every thread streams its slice of the weights with
// its loads issued up-front (unrolled), so a call costs ~2 memory latencies:
//   NN (forward, B = W[K,N] row-major): CTA = 16 columns; 4 float4 column
//       threads x 64 k-lanes; shuffle + smem reduction over the k-lanes.
//   NT (input gradient, B = W[N,K] row-major): one warp per output column,
//       lanes stride K with float4; warp-shuffle reduction.
// Replaces the M=B `x @ kernel` calls inside RSSM.obs_step (reference
// nets.py:99-117) and their input gradients.  (r01 profile: the first, chunked
// version of these kernels took ~20 us per call = 28% of the train step.)
#include "dd_common.cuh"

namespace dd {

constexpr int SK_SMEM_MAX = 200 * 1024;

template <int MT>
__global__ void __launch_bounds__(256)
skinny_nn_kernel(const float* __restrict__ A, int lda, const float* __restrict__ B, int ldb,
                 float* __restrict__ C, int ldc, int M, int N, int K, int KC, float beta,
                 const float* __restrict__ bias) {
  pdl_prologue();
  extern __shared__ __align__(16) float sk_smem[];
  float (*As)[MT + 4] = reinterpret_cast<float (*)[MT + 4]>(sk_smem);       // [KC][MT+4]
  float (*red)[MT][16] = reinterpret_cast<float (*)[MT][16]>(sk_smem);      // [8][MT][16]
  const int tid = threadIdx.x, c4 = tid & 3, kl = tid >> 2;                // 4 col threads x 64 k-lanes
  const int n0 = blockIdx.x * 16 + c4 * 4;
  const bool nvec = (n0 + 3 < N) && ((ldb & 3) == 0) &&
                    ((reinterpret_cast<uintptr_t>(B) & 15) == 0);
  float acc[MT][4];
#pragma unroll
  for (int m = 0; m < MT; ++m) acc[m][0] = acc[m][1] = acc[m][2] = acc[m][3] = 0.f;

  for (int k0 = 0; k0 < K; k0 += KC) {
    const int kc = min(KC, K - k0);
    __syncthreads();
    const int tot = MT * kc;
#pragma unroll 8
    for (int e = tid; e < tot; e += 256) {
      const int m = e / kc, k = e - m * kc;                                 // coalesced along k
      As[k][m] = m < M ? A[(size_t)m * lda + k0 + k] : 0.f;
    }
    __syncthreads();
#pragma unroll 4
    for (int k = kl; k < kc; k += 64) {
      float4 b;
      const float* bp = B + (size_t)(k0 + k) * ldb + n0;
      if (nvec) {
        b = *reinterpret_cast<const float4*>(bp);
      } else {
        b.x = n0 + 0 < N ? bp[0] : 0.f; b.y = n0 + 1 < N ? bp[1] : 0.f;
        b.z = n0 + 2 < N ? bp[2] : 0.f; b.w = n0 + 3 < N ? bp[3] : 0.f;
      }
#pragma unroll
      for (int m4 = 0; m4 < MT / 4; ++m4) {
        const float4 a = *reinterpret_cast<const float4*>(&As[k][m4 * 4]);
        const float av[4] = {a.x, a.y, a.z, a.w};
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          acc[m4 * 4 + i][0] = fmaf(av[i], b.x, acc[m4 * 4 + i][0]);
          acc[m4 * 4 + i][1] = fmaf(av[i], b.y, acc[m4 * 4 + i][1]);
          acc[m4 * 4 + i][2] = fmaf(av[i], b.z, acc[m4 * 4 + i][2]);
          acc[m4 * 4 + i][3] = fmaf(av[i], b.w, acc[m4 * 4 + i][3]);
        }
      }
    }
  }
  // stage 1: the 8 k-lanes of a warp (lane bits 2,3,4)
#pragma unroll
  for (int m = 0; m < MT; ++m)
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      float v = acc[m][j];
      v += __shfl_xor_sync(0xffffffffu, v, 4);
      v += __shfl_xor_sync(0xffffffffu, v, 8);
      v += __shfl_xor_sync(0xffffffffu, v, 16);
      acc[m][j] = v;
    }
  const int warp = tid >> 5, lane = tid & 31;
  __syncthreads();                                   // everyone is done reading As
  if (lane < 4) {
#pragma unroll
    for (int m = 0; m < MT; ++m)
#pragma unroll
      for (int j = 0; j < 4; ++j) red[warp][m][c4 * 4 + j] = acc[m][j];
  }
  __syncthreads();
  // stage 2: the 8 warps
  for (int e = tid; e < MT * 16; e += 256) {
    const int m = e / 16, c = e % 16, n = blockIdx.x * 16 + c;
    if (m < M && n < N) {
      float v = 0.f;
#pragma unroll
      for (int w = 0; w < 8; ++w) v += red[w][m][c];
      if (bias != nullptr) v += bias[n];
      float* cp = &C[(size_t)m * ldc + n];
      *cp = (beta != 0.f) ? v + beta * (*cp) : v;
    }
  }
}

template <int MT>
__global__ void __launch_bounds__(256)
skinny_nt_kernel(const float* __restrict__ A, int lda, const float* __restrict__ B, int ldb,
                 float* __restrict__ C, int ldc, int M, int N, int K, int KC, float beta,
                 const float* __restrict__ bias) {
  pdl_prologue();
  extern __shared__ __align__(16) float sk_smem[];                          // [MT][KC + 4]
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int n = blockIdx.x * 8 + warp;
  const int lds = KC + 4;
  const bool kvec = ((ldb & 3) == 0) && ((reinterpret_cast<uintptr_t>(B) & 15) == 0);
  float acc[MT];
#pragma unroll
  for (int m = 0; m < MT; ++m) acc[m] = 0.f;
  for (int k0 = 0; k0 < K; k0 += KC) {
    const int kc = min(KC, K - k0);
    __syncthreads();
    const int tot = MT * kc;
#pragma unroll 8
    for (int e = tid; e < tot; e += 256) {
      const int m = e / kc, k = e - m * kc;
      sk_smem[m * lds + k] = m < M ? A[(size_t)m * lda + k0 + k] : 0.f;
    }
    // zero the float4 tail so vector reads past kc see zeros
    if (tid < MT * 4) sk_smem[(tid >> 2) * lds + kc + (tid & 3)] = 0.f;
    __syncthreads();
    if (n < N) {
#pragma unroll 4
      for (int k = lane * 4; k < kc; k += 128) {
        const float* bp = B + (size_t)n * ldb + k0 + k;
        float4 b;
        if (kvec && k + 3 < kc) {
          b = *reinterpret_cast<const float4*>(bp);
        } else {
          b.x = k + 0 < kc ? bp[0] : 0.f; b.y = k + 1 < kc ? bp[1] : 0.f;
          b.z = k + 2 < kc ? bp[2] : 0.f; b.w = k + 3 < kc ? bp[3] : 0.f;
        }
#pragma unroll
        for (int m = 0; m < MT; ++m) {
          const float4 a = *reinterpret_cast<const float4*>(&sk_smem[m * lds + k]);
          acc[m] = fmaf(a.x, b.x, fmaf(a.y, b.y, fmaf(a.z, b.z, fmaf(a.w, b.w, acc[m]))));
        }
      }
    }
  }
#pragma unroll
  for (int m = 0; m < MT; ++m) acc[m] = warp_sum(acc[m]);
  if (lane == 0 && n < N) {
#pragma unroll
    for (int m = 0; m < MT; ++m) {
      if (m < M) {
        float v = acc[m];
        if (bias != nullptr) v += bias[n];
        float* cp = &C[(size_t)m * ldc + n];
        *cp = (beta != 0.f) ? v + beta * (*cp) : v;
      }
    }
  }
}

template <typename Kern>
static void allow_big_smem(Kern kern) {
  cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, SK_SMEM_MAX);
}

// 1 = handled, 0 = not eligible
int gemm_skinny_try(const float* A, int lda, int ta, const float* B, int ldb, int tb, float* C,
                    int ldc, int M, int N, int K, float beta, const float* bias,
                    cudaStream_t st) {
  if (ta || M > 32 || N < 16 || K < 64) return 0;
  static bool once = false;
  if (!once) {
    allow_big_smem(skinny_nn_kernel<16>);
    allow_big_smem(skinny_nn_kernel<32>);
    allow_big_smem(skinny_nt_kernel<16>);
    allow_big_smem(skinny_nt_kernel<32>);
    once = true;
  }
  const int MT = M <= 16 ? 16 : 32;
  if (!tb) {
    const int per_k = (MT + 4) * 4;
    int KC = min((K + 3) / 4 * 4, SK_SMEM_MAX / per_k / 64 * 64);
    const int red_bytes = 8 * MT * 16 * 4;
    const int smem = max(KC * per_k, red_bytes);
    const int grid = (N + 15) / 16;
    if (MT == 16) launch_k(skinny_nn_kernel<16>, grid, 256, smem, st, A, lda, B, ldb, C, ldc, M, N, K, KC, beta, bias);
    else launch_k(skinny_nn_kernel<32>, grid, 256, smem, st, A, lda, B, ldb, C, ldc, M, N, K, KC, beta, bias);
  } else {
    int KC = min((K + 3) / 4 * 4, (SK_SMEM_MAX / (MT * 4) - 4) / 128 * 128);
    const int smem = MT * (KC + 4) * 4;
    const int grid = (N + 7) / 8;
    if (MT == 16) launch_k(skinny_nt_kernel<16>, grid, 256, smem, st, A, lda, B, ldb, C, ldc, M, N, K, KC, beta, bias);
    else launch_k(skinny_nt_kernel<32>, grid, 256, smem, st, A, lda, B, ldb, C, ldc, M, N, K, KC, beta, bias);
  }
  return launched("dd_gemm_skinny") == 0 ? 1 : -1;
}

}  // namespace dd
