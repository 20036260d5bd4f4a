// Counter-based noise for the categorical (Gumbel) and action (normal) draws.
// Philox4x32-10 (Salmon et al., SC'11).  The reference draws from TF's global
// stateful stream (tfutils.py:374 tf.random.categorical, seed=None), which is
// irreproducible outside TF; parity is therefore defined on injected noise and
// this generator is mirrored bit-for-bit (up to libm ulps) by oracle/philox.py.
//
// Layout: element e uses block = e / 4, lane = e % 4;
//   counter = (block_lo, block_hi, stream_id, step), key = (seed_lo, seed_hi);
//   u = ((x >> 8) + 0.5) * 2^-24  in (0, 1).
#include "dd_common.cuh"

namespace dd {

__device__ __forceinline__ void philox4x32_10(uint32_t c[4], uint32_t k0, uint32_t k1) {
  const uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    const uint32_t hi0 = __umulhi(M0, c[0]), lo0 = M0 * c[0];
    const uint32_t hi1 = __umulhi(M1, c[2]), lo1 = M1 * c[2];
    const uint32_t n0 = hi1 ^ c[1] ^ k0, n1 = lo1, n2 = hi0 ^ c[3] ^ k1, n3 = lo0;
    c[0] = n0; c[1] = n1; c[2] = n2; c[3] = n3;
    k0 += W0;
    k1 += W1;
  }
}

__device__ __forceinline__ float u01(uint32_t x) {
  return ((float)(x >> 8) + 0.5f) * 5.9604644775390625e-8f;  // 2^-24
}

__global__ void fill_noise_kernel(float* __restrict__ buf, int64_t n, uint64_t seed,
                                  uint32_t stream_id, const int64_t* __restrict__ counters,
                                  int ci, int normal) {
  pdl_prologue();
  const uint32_t step = (uint32_t)(counters[ci] & 0xffffffffll);
  const int64_t nb = (n + 3) / 4;
  for (int64_t b = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; b < nb;
       b += (int64_t)gridDim.x * blockDim.x) {
    uint32_t c[4] = {(uint32_t)(b & 0xffffffffll), (uint32_t)(b >> 32), stream_id, step};
    philox4x32_10(c, (uint32_t)(seed & 0xffffffffull), (uint32_t)(seed >> 32));
    float o[4];
    if (normal) {
      const float two_pi = 6.283185307179586f;
      const float r0 = sqrtf(-2.f * logf(u01(c[0]))), a0 = two_pi * u01(c[1]);
      const float r1 = sqrtf(-2.f * logf(u01(c[2]))), a1 = two_pi * u01(c[3]);
      o[0] = r0 * cosf(a0); o[1] = r0 * sinf(a0);
      o[2] = r1 * cosf(a1); o[3] = r1 * sinf(a1);
    } else {
#pragma unroll
      for (int i = 0; i < 4; ++i) o[i] = -logf(-logf(u01(c[i])));
    }
#pragma unroll
    for (int i = 0; i < 4; ++i)
      if (4 * b + i < n) buf[4 * b + i] = o[i];
  }
}

}  // namespace dd

using namespace dd;

extern "C" int dd_fill_gumbel(float* buf, int64_t n, uint64_t seed, uint32_t stream_id,
                              const int64_t* counters, int ci, dd_stream_t stream) {
  if (n <= 0) return 0;
  launch_k(fill_noise_kernel, grid_for((n + 3) / 4, 256), 256, 0, S(stream), buf, n, seed, stream_id,
                                                                       counters, ci, 0);
  return launched("dd_fill_gumbel");
}

extern "C" int dd_fill_normal(float* buf, int64_t n, uint64_t seed, uint32_t stream_id,
                              const int64_t* counters, int ci, dd_stream_t stream) {
  if (n <= 0) return 0;
  launch_k(fill_noise_kernel, grid_for((n + 3) / 4, 256), 256, 0, S(stream), buf, n, seed, stream_id,
                                                                       counters, ci, 1);
  return launched("dd_fill_normal");
}
