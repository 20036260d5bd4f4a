// Row-wise fused operators of the RSSM cell: LayerNorm(+elu) forward/backward,
// GRU gates, the categorical latent (unimix softmax, Gumbel-argmax draw,
// straight-through backward), KL balance, and the is_first masking of the
// observe scan.  Reference: embodied/agents/dreamerv2plus/nets.py:99-183,
// :585-602; tfutils.py:359-387.  All HBM-bound: one pass over the row, fp32.
#include "dd_common.cuh"

namespace dd {

// ---------------------------------------------------------------------------
// LayerNorm + optional elu.  One CTA per row; the row is staged in shared
// memory so the mean / centred variance / normalise passes touch HBM once.
// ---------------------------------------------------------------------------
__global__ void ln_fwd_kernel(const float* __restrict__ Z, int ldz,
                              const float* __restrict__ scale,
                              const float* __restrict__ bias, float* __restrict__ H, int ldh,
                              float* __restrict__ stats, int M, int N, int act) {
  pdl_prologue();
  extern __shared__ float row[];
  __shared__ float red[33];
  for (int m = blockIdx.x; m < M; m += gridDim.x) {
    const float* z = Z + (size_t)m * ldz;
    float s = 0.f;
    for (int j = threadIdx.x; j < N; j += blockDim.x) {
      const float v = z[j];
      row[j] = v;
      s += v;
    }
    const float mean = block_sum(s, red) / N;
    float q = 0.f;
    for (int j = threadIdx.x; j < N; j += blockDim.x) {
      const float d = row[j] - mean;
      q += d * d;
    }
    const float var = block_sum(q, red) / N;
    const float rstd = rsqrtf(var + DD_LN_EPS);
    float* h = H + (size_t)m * ldh;
    for (int j = threadIdx.x; j < N; j += blockDim.x) {
      float y = (row[j] - mean) * rstd * scale[j] + bias[j];
      if (act) y = y > 0.f ? y : expm1f(y);
      h[j] = y;
    }
    if (threadIdx.x == 0) {
      stats[2 * m] = mean;
      stats[2 * m + 1] = rstd;
    }
    __syncthreads();
  }
}

// Backward.  Each CTA walks a strided set of rows, keeps the per-column
// dscale/dbias partial sums in registers and issues one atomicAdd per column at
// the end (M/gridDim-fold fewer atomics than per-row).
template <int PER, int BLK>
__global__ void __launch_bounds__(BLK) ln_bwd_kernel(const float* __restrict__ dH, int lddh,
                              const float* __restrict__ Z, int ldz,
                              const float* __restrict__ stats,
                              const float* __restrict__ scale,
                              const float* __restrict__ bias, float* __restrict__ dZ,
                              int lddz, float* __restrict__ dscale,
                              float* __restrict__ dbias, int M, int N, int act) {
  pdl_prologue();
  __shared__ float red[33];
  float gs[PER], gb[PER];
#pragma unroll
  for (int i = 0; i < PER; ++i) gs[i] = gb[i] = 0.f;
  for (int m = blockIdx.x; m < M; m += gridDim.x) {
    const float mean = stats[2 * m], rstd = stats[2 * m + 1];
    const float* z = Z + (size_t)m * ldz;
    const float* dh = dH + (size_t)m * lddh;
    float xh[PER], dx[PER];
    float s1 = 0.f, s2 = 0.f;
#pragma unroll
    for (int i = 0; i < PER; ++i) {
      const int j = threadIdx.x + i * blockDim.x;
      xh[i] = 0.f;
      dx[i] = 0.f;
      if (j < N) {
        const float x = (z[j] - mean) * rstd;
        const float sc = scale[j];
        float dy = dh[j];
        if (act) {
          const float y = x * sc + bias[j];
          if (y <= 0.f) dy *= expf(y);
        }
        gs[i] += dy * x;
        gb[i] += dy;
        xh[i] = x;
        dx[i] = dy * sc;
        s1 += dx[i];
        s2 += dx[i] * x;
      }
    }
    const float m1 = block_sum(s1, red) / N;
    const float m2 = block_sum(s2, red) / N;
    float* dz = dZ + (size_t)m * lddz;
#pragma unroll
    for (int i = 0; i < PER; ++i) {
      const int j = threadIdx.x + i * blockDim.x;
      if (j < N) dz[j] = rstd * (dx[i] - m1 - xh[i] * m2);
    }
  }
  if (dscale != nullptr) {
#pragma unroll
    for (int i = 0; i < PER; ++i) {
      const int j = threadIdx.x + i * blockDim.x;
      if (j < N) {
        atomicAdd(&dscale[j], gs[i]);
        atomicAdd(&dbias[j], gb[i]);
      }
    }
  }
}

// ---------------------------------------------------------------------------
// GRU gates (nets.py:155-160)
// ---------------------------------------------------------------------------
__global__ void gru_fwd_kernel(const float* __restrict__ Y, int ldy,
                               const float* __restrict__ prev, int ldp,
                               float* __restrict__ out, int ldo, int M, int D) {
  pdl_prologue();
  const int64_t n = (int64_t)M * D;
  for (int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; e < n;
       e += (int64_t)gridDim.x * blockDim.x) {
    const int m = e / D, j = e % D;
    const float* y = Y + (size_t)m * ldy;
    const float r = sigmoidf_(y[j]);
    const float c = tanhf(r * y[D + j]);
    const float u = sigmoidf_(y[2 * D + j] - 1.f);
    out[(size_t)m * ldo + j] = u * c + (1.f - u) * prev[(size_t)m * ldp + j];
  }
}

__global__ void gru_bwd_kernel(const float* __restrict__ dOut, int lddo,
                               const float* __restrict__ Y, int ldy,
                               const float* __restrict__ prev, int ldp,
                               float* __restrict__ dY, int lddy, float* __restrict__ dPrev,
                               int lddp, int M, int D, int acc) {
  pdl_prologue();
  const int64_t n = (int64_t)M * D;
  for (int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; e < n;
       e += (int64_t)gridDim.x * blockDim.x) {
    const int m = e / D, j = e % D;
    const float* y = Y + (size_t)m * ldy;
    const float yc = y[D + j];
    const float r = sigmoidf_(y[j]);
    const float c = tanhf(r * yc);
    const float u = sigmoidf_(y[2 * D + j] - 1.f);
    const float p = prev[(size_t)m * ldp + j];
    const float g = dOut[(size_t)m * lddo + j];
    const float dpre = g * u * (1.f - c * c);
    float* dy = dY + (size_t)m * lddy;
    dy[j] = dpre * yc * r * (1.f - r);
    dy[D + j] = dpre * r;
    dy[2 * D + j] = g * (c - p) * u * (1.f - u);
    float* dp = &dPrev[(size_t)m * lddp + j];
    const float v = g * (1.f - u);
    *dp = acc ? *dp + v : v;
  }
}

// ---------------------------------------------------------------------------
// Categorical latent.  One warp per (row, group); C <= 64 classes.
// ---------------------------------------------------------------------------
__global__ void stats_fwd_kernel(const float* __restrict__ Z, int ldz,
                                 const float* __restrict__ G, int ldg,
                                 float* __restrict__ logit, int ldl,
                                 float* __restrict__ stoch, int lds,
                                 int32_t* __restrict__ idx, int M, int S, int C, float mix,
                                 int mode) {
  pdl_prologue();
  const int lane = threadIdx.x & 31;
  const int wpb = blockDim.x >> 5;
  const int64_t ngroups = (int64_t)M * S;
  for (int64_t gidx = blockIdx.x * (int64_t)wpb + (threadIdx.x >> 5); gidx < ngroups;
       gidx += (int64_t)gridDim.x * wpb) {
    const int m = gidx / S, s = gidx % S;
    const float* z = Z + (size_t)m * ldz + s * C;
    const int c0 = lane, c1 = lane + 32;
    const float z0 = c0 < C ? z[c0] : -INFINITY;
    const float z1 = c1 < C ? z[c1] : -INFINITY;
    const float mx = warp_max(fmaxf(z0, z1));
    const float e0 = c0 < C ? expf(z0 - mx) : 0.f;
    const float e1 = c1 < C ? expf(z1 - mx) : 0.f;
    const float inv = 1.f / warp_sum(e0 + e1);
    const float u = mix / C;
    const float l0 = logf((1.f - mix) * e0 * inv + u);
    const float l1 = logf((1.f - mix) * e1 * inv + u);
    float* lg = logit + (size_t)m * ldl + s * C;
    if (c0 < C) lg[c0] = l0;
    if (c1 < C) lg[c1] = l1;
    if (mode == 2) continue;
    float k0 = l0, k1 = l1;
    if (mode == 0) {
      const float* g = G + (size_t)m * ldg + s * C;
      if (c0 < C) k0 += g[c0];
      if (c1 < C) k1 += g[c1];
    }
    // argmax with first-index tie break (torch.argmax / tf.argmax semantics)
    float bv = c0 < C ? k0 : -INFINITY;
    int bi = c0 < C ? c0 : 0x7fffffff;
    if (c1 < C && k1 > bv) { bv = k1; bi = c1; }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const float ov = __shfl_xor_sync(0xffffffffu, bv, o);
      const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
      if (ov > bv || (ov == bv && oi < bi)) { bv = ov; bi = oi; }
    }
    float* st = stoch + (size_t)m * lds + s * C;
    if (c0 < C) st[c0] = (c0 == bi) ? 1.f : 0.f;
    if (c1 < C) st[c1] = (c1 == bi) ? 1.f : 0.f;
    if (idx != nullptr && lane == 0) idx[(size_t)m * S + s] = bi;
  }
}

__global__ void stats_bwd_kernel(const float* __restrict__ dStoch, int ldds,
                                 const float* __restrict__ dLogit, int lddl,
                                 const float* __restrict__ logit, int ldl,
                                 float* __restrict__ dZ, int lddz, int M, int S, int C,
                                 float mix) {
  pdl_prologue();
  const int lane = threadIdx.x & 31;
  const int wpb = blockDim.x >> 5;
  const int64_t ngroups = (int64_t)M * S;
  const float u = mix / C;
  for (int64_t gidx = blockIdx.x * (int64_t)wpb + (threadIdx.x >> 5); gidx < ngroups;
       gidx += (int64_t)gridDim.x * wpb) {
    const int m = gidx / S, s = gidx % S;
    const float* lg = logit + (size_t)m * ldl + s * C;
    float q[2], dl[2];
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      const int c = lane + 32 * i;
      q[i] = c < C ? expf(lg[c]) : 0.f;
      dl[i] = 0.f;
    }
    if (dStoch != nullptr) {
      const float* ds = dStoch + (size_t)m * ldds + s * C;
      float d[2], dot = 0.f;
#pragma unroll
      for (int i = 0; i < 2; ++i) {
        const int c = lane + 32 * i;
        d[i] = c < C ? ds[c] : 0.f;
        dot += q[i] * d[i];
      }
      dot = warp_sum(dot);
#pragma unroll
      for (int i = 0; i < 2; ++i) dl[i] = q[i] * (d[i] - dot);
    }
    if (dLogit != nullptr) {
      const float* d = dLogit + (size_t)m * lddl + s * C;
#pragma unroll
      for (int i = 0; i < 2; ++i) {
        const int c = lane + 32 * i;
        if (c < C) dl[i] += d[c];
      }
    }
    float p[2], dp[2], dot = 0.f;
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      const int c = lane + 32 * i;
      p[i] = c < C ? (q[i] - u) / (1.f - mix) : 0.f;
      dp[i] = c < C ? (1.f - mix) * dl[i] / q[i] : 0.f;
      dot += p[i] * dp[i];
    }
    dot = warp_sum(dot);
    float* dz = dZ + (size_t)m * lddz + s * C;
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      const int c = lane + 32 * i;
      if (c < C) dz[c] = p[i] * (dp[i] - dot);
    }
  }
}

// log_softmax of one group held two-per-lane
__device__ __forceinline__ void group_lsm(const float* x, int C, int lane, float out[2]) {
  float v[2];
#pragma unroll
  for (int i = 0; i < 2; ++i) {
    const int c = lane + 32 * i;
    v[i] = c < C ? x[c] : -INFINITY;
  }
  const float mx = warp_max(fmaxf(v[0], v[1]));
  float e = 0.f;
#pragma unroll
  for (int i = 0; i < 2; ++i) e += (lane + 32 * i) < C ? expf(v[i] - mx) : 0.f;
  const float lse = mx + logf(warp_sum(e));
#pragma unroll
  for (int i = 0; i < 2; ++i) out[i] = v[i] - lse;
}

// One CTA per row: warps stride over the S groups, then a block reduction.
__global__ void kl_fwd_kernel(const float* __restrict__ post, int ldp,
                              const float* __restrict__ prior, int ldq, int M, int S, int C,
                              float* __restrict__ kl_row, float* __restrict__ ent_post,
                              float* __restrict__ ent_prior) {
  pdl_prologue();
  __shared__ float red[33];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
  for (int m = blockIdx.x; m < M; m += gridDim.x) {
    float kl = 0.f, ea = 0.f, eb = 0.f;
    for (int s = warp; s < S; s += nw) {
      float a[2], b[2];
      group_lsm(post + (size_t)m * ldp + s * C, C, lane, a);
      group_lsm(prior + (size_t)m * ldq + s * C, C, lane, b);
#pragma unroll
      for (int i = 0; i < 2; ++i) {
        if (lane + 32 * i < C) {
          const float pa = expf(a[i]), pb = expf(b[i]);
          kl += pa * (a[i] - b[i]);
          ea -= pa * a[i];
          eb -= pb * b[i];
        }
      }
    }
    kl = block_sum(kl, red);
    ea = block_sum(ea, red);
    eb = block_sum(eb, red);
    if (threadIdx.x == 0) {
      kl_row[m] = kl;
      ent_post[m] = ea;
      ent_prior[m] = eb;
    }
  }
}

__global__ void kl_bwd_kernel(const float* __restrict__ post, int ldp,
                              const float* __restrict__ prior, int ldq, int M, int S, int C,
                              float bal, float coef, const float* __restrict__ scale,
                              float* __restrict__ dpost, int lddp, float* __restrict__ dprior,
                              int lddq) {
  pdl_prologue();
  const int lane = threadIdx.x & 31;
  const int wpb = blockDim.x >> 5;
  const int64_t ngroups = (int64_t)M * S;
  const float k = coef * scale[0];
  for (int64_t gidx = blockIdx.x * (int64_t)wpb + (threadIdx.x >> 5); gidx < ngroups;
       gidx += (int64_t)gridDim.x * wpb) {
    const int m = gidx / S, s = gidx % S;
    float a[2], b[2], pa[2], pb[2];
    group_lsm(post + (size_t)m * ldp + s * C, C, lane, a);
    group_lsm(prior + (size_t)m * ldq + s * C, C, lane, b);
    float kl = 0.f;
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      const bool ok = lane + 32 * i < C;
      pa[i] = ok ? expf(a[i]) : 0.f;
      pb[i] = ok ? expf(b[i]) : 0.f;
      if (ok) kl += pa[i] * (a[i] - b[i]);
    }
    kl = warp_sum(kl);
    float* dp = dpost + (size_t)m * lddp + s * C;
    float* dq = dprior + (size_t)m * lddq + s * C;
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      const int c = lane + 32 * i;
      if (c < C) {
        dp[c] = k * (1.f - bal) * pa[i] * ((a[i] - b[i]) - kl);
        dq[c] = k * bal * (pb[i] - pa[i]);
      }
    }
  }
}

// ---------------------------------------------------------------------------
// is_first masking of the scan carry (nets.py:100-107)
// ---------------------------------------------------------------------------
__global__ void mask_fwd_kernel(const float* __restrict__ pd, int ldpd,
                                const float* __restrict__ ps, int ldps,
                                const float* __restrict__ act, int lda,
                                const float* __restrict__ first,
                                const float* __restrict__ init_d,
                                const float* __restrict__ init_s, float* __restrict__ od,
                                int ldod, float* __restrict__ os, int ldos,
                                float* __restrict__ oa, int ldoa, int M, int D, int SC, int A) {
  pdl_prologue();
  const int W = D + SC + A;
  const int64_t n = (int64_t)M * W;
  for (int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; e < n;
       e += (int64_t)gridDim.x * blockDim.x) {
    const int m = e / W;
    int j = e % W;
    const float f = first[m], k = 1.f - f;
    if (j < D) {
      od[(size_t)m * ldod + j] = pd[(size_t)m * ldpd + j] * k + init_d[j] * f;
    } else if ((j -= D) < SC) {
      os[(size_t)m * ldos + j] = ps[(size_t)m * ldps + j] * k + init_s[j] * f;
    } else {
      j -= SC;
      oa[(size_t)m * ldoa + j] = act[(size_t)m * lda + j] * k;
    }
  }
}

__global__ void mask_bwd_kernel(const float* __restrict__ dd_, int lddd,
                                const float* __restrict__ ds, int ldds,
                                const float* __restrict__ first, float* __restrict__ dpd,
                                int lddpd, float* __restrict__ dps, int lddps,
                                float* __restrict__ dinit, int M, int D, int SC) {
  pdl_prologue();
  const int W = D + SC;
  for (int j = blockIdx.x * blockDim.x + threadIdx.x; j < W; j += gridDim.x * blockDim.x) {
    if (j < D) {
      float acc = 0.f;
      for (int m = 0; m < M; ++m) {
        const float f = first[m];
        const float g = dd_[(size_t)m * lddd + j];
        acc += g * f;
        if (dpd != nullptr) dpd[(size_t)m * lddpd + j] += g * (1.f - f);
      }
      dinit[j] += acc;
    } else if (dps != nullptr) {
      const int c = j - D;
      for (int m = 0; m < M; ++m)
        dps[(size_t)m * lddps + c] += ds[(size_t)m * ldds + c] * (1.f - first[m]);
    }
  }
}

// Rows of up to 512 values (the per-pixel channel LayerNorm of the conv stacks with millions of
// rows, and every 512-wide MLP layer): one warp per row, the row held in registers (<= 16 values
// per lane), warp-shuffle reductions, no block barriers.  Default limit 128 columns: measured on
// the a1 step (r01), the warp kernels at N = 512 were SLOWER than block-per-row (fewer warps in
// flight per SM: imagine +0.5 ms, actor_update +2 ms), so wider rows use them only on request
// (dd_set_ln_warp_max(512)).
int g_ln_warp_max = 128;

template <int PER>
__global__ void __launch_bounds__(256) ln_fwd_warp_kernel(
    const float* __restrict__ Z, int ldz, const float* __restrict__ scale,
    const float* __restrict__ bias, float* __restrict__ H, int ldh, float* __restrict__ stats,
    int M, int N, int act) {
  pdl_prologue();
  const int lane = threadIdx.x & 31, wpb = blockDim.x >> 5;
  float sc[PER], bi[PER];
#pragma unroll
  for (int i = 0; i < PER; ++i) {
    const int j = lane + 32 * i;
    sc[i] = j < N ? scale[j] : 0.f;
    bi[i] = j < N ? bias[j] : 0.f;
  }
  for (int64_t m = blockIdx.x * (int64_t)wpb + (threadIdx.x >> 5); m < M;
       m += (int64_t)gridDim.x * wpb) {
    const float* z = Z + (size_t)m * ldz;
    float v[PER], s = 0.f;
#pragma unroll
    for (int i = 0; i < PER; ++i) {
      const int j = lane + 32 * i;
      v[i] = j < N ? z[j] : 0.f;
      s += v[i];
    }
    const float mean = warp_sum(s) / N;
    float q = 0.f;
#pragma unroll
    for (int i = 0; i < PER; ++i) {
      const float d = (lane + 32 * i) < N ? v[i] - mean : 0.f;
      q += d * d;
    }
    const float rstd = rsqrtf(warp_sum(q) / N + DD_LN_EPS);
    float* h = H + (size_t)m * ldh;
#pragma unroll
    for (int i = 0; i < PER; ++i) {
      const int j = lane + 32 * i;
      if (j < N) {
        float y = (v[i] - mean) * rstd * sc[i] + bi[i];
        if (act) y = y > 0.f ? y : expm1f(y);
        h[j] = y;
      }
    }
    if (lane == 0) {
      stats[2 * m] = mean;
      stats[2 * m + 1] = rstd;
    }
  }
}

template <int PER>
__global__ void __launch_bounds__(256) ln_bwd_warp_kernel(
    const float* __restrict__ dH, int lddh, const float* __restrict__ Z, int ldz,
    const float* __restrict__ stats, const float* __restrict__ scale,
    const float* __restrict__ bias, float* __restrict__ dZ, int lddz,
    float* __restrict__ dscale, float* __restrict__ dbias, int M, int N, int act) {
  pdl_prologue();
  __shared__ float sh_s[8][32 * PER], sh_b[8][32 * PER];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, wpb = blockDim.x >> 5;
  float sc[PER], bi[PER], gs[PER], gb[PER];
#pragma unroll
  for (int i = 0; i < PER; ++i) {
    const int j = lane + 32 * i;
    sc[i] = j < N ? scale[j] : 0.f;
    bi[i] = j < N ? bias[j] : 0.f;
    gs[i] = gb[i] = 0.f;
  }
  for (int64_t m = blockIdx.x * (int64_t)wpb + warp; m < M; m += (int64_t)gridDim.x * wpb) {
    const float mean = stats[2 * m], rstd = stats[2 * m + 1];
    const float* z = Z + (size_t)m * ldz;
    const float* dh = dH + (size_t)m * lddh;
    float xh[PER], dx[PER], s1 = 0.f, s2 = 0.f;
#pragma unroll
    for (int i = 0; i < PER; ++i) {
      const int j = lane + 32 * i;
      xh[i] = dx[i] = 0.f;
      if (j < N) {
        const float x = (z[j] - mean) * rstd;
        float dy = dh[j];
        if (act) {
          const float y = x * sc[i] + bi[i];
          if (y <= 0.f) dy *= expf(y);
        }
        gs[i] += dy * x;
        gb[i] += dy;
        xh[i] = x;
        dx[i] = dy * sc[i];
        s1 += dx[i];
        s2 += dx[i] * x;
      }
    }
    const float m1 = warp_sum(s1) / N, m2 = warp_sum(s2) / N;
    float* dz = dZ + (size_t)m * lddz;
#pragma unroll
    for (int i = 0; i < PER; ++i) {
      const int j = lane + 32 * i;
      if (j < N) dz[j] = rstd * (dx[i] - m1 - xh[i] * m2);
    }
  }
  if (dscale == nullptr) return;
#pragma unroll
  for (int i = 0; i < PER; ++i) {
    sh_s[warp][lane + 32 * i] = gs[i];
    sh_b[warp][lane + 32 * i] = gb[i];
  }
  __syncthreads();
  for (int j = threadIdx.x; j < N; j += blockDim.x) {
    float a = 0.f, b = 0.f;
    for (int w = 0; w < wpb; ++w) {
      a += sh_s[w][j];
      b += sh_b[w][j];
    }
    atomicAdd(&dscale[j], a);
    atomicAdd(&dbias[j], b);
  }
}

// ---- EXPERIMENTAL (round 2, not yet run on a GPU; dd_set_ln_vec) -----------------------------
// r01 measured the block-per-row kernels at 26-29 % of the HBM copy peak for M ~ 15K rows and
// latency bound (7-10 us) at M = 1024 (profiles/r01_gemm_shape_times.txt): three block barriers
// per row, scalar 4-byte accesses, ~32 KB in flight per SM.  Variant for N % 128 == 0, N <= 512:
// one warp per row, float4 accesses (lane owns columns 4*lane + 128*i), no block barrier, register
// budget capped so 3-4 CTAs of 8 warps are resident per SM.
int g_ln_vec = 1;

template <int V>
__global__ void __launch_bounds__(256, 4) ln_fwd_vec_kernel(
    const float* __restrict__ Z, int ldz, const float* __restrict__ scale,
    const float* __restrict__ bias, float* __restrict__ H, int ldh, float* __restrict__ stats,
    int M, int N, int act) {
  pdl_prologue();
  const int lane = threadIdx.x & 31, wpb = blockDim.x >> 5;
  const float4* sc4 = reinterpret_cast<const float4*>(scale);
  const float4* bi4 = reinterpret_cast<const float4*>(bias);
  for (int64_t m = blockIdx.x * (int64_t)wpb + (threadIdx.x >> 5); m < M;
       m += (int64_t)gridDim.x * wpb) {
    const float4* z4 = reinterpret_cast<const float4*>(Z + (size_t)m * ldz);
    float4 v[V];
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < V; ++i) {
      v[i] = z4[lane + 32 * i];
      s += (v[i].x + v[i].y) + (v[i].z + v[i].w);
    }
    const float mean = warp_sum(s) / N;
    float q = 0.f;
#pragma unroll
    for (int i = 0; i < V; ++i) {
      const float a = v[i].x - mean, b = v[i].y - mean, c = v[i].z - mean, d = v[i].w - mean;
      q += (a * a + b * b) + (c * c + d * d);
    }
    const float rstd = rsqrtf(warp_sum(q) / N + DD_LN_EPS);
    float4* h4 = reinterpret_cast<float4*>(H + (size_t)m * ldh);
#pragma unroll
    for (int i = 0; i < V; ++i) {
      const float4 sc = sc4[lane + 32 * i], bi = bi4[lane + 32 * i];
      float4 y;
      y.x = (v[i].x - mean) * rstd * sc.x + bi.x;
      y.y = (v[i].y - mean) * rstd * sc.y + bi.y;
      y.z = (v[i].z - mean) * rstd * sc.z + bi.z;
      y.w = (v[i].w - mean) * rstd * sc.w + bi.w;
      if (act) {
        y.x = y.x > 0.f ? y.x : expm1f(y.x);
        y.y = y.y > 0.f ? y.y : expm1f(y.y);
        y.z = y.z > 0.f ? y.z : expm1f(y.z);
        y.w = y.w > 0.f ? y.w : expm1f(y.w);
      }
      h4[lane + 32 * i] = y;
    }
    if (lane == 0) {
      stats[2 * m] = mean;
      stats[2 * m + 1] = rstd;
    }
  }
}

template <int V>
__global__ void __launch_bounds__(256, 2) ln_bwd_vec_kernel(
    const float* __restrict__ dH, int lddh, const float* __restrict__ Z, int ldz,
    const float* __restrict__ stats, const float* __restrict__ scale,
    const float* __restrict__ bias, float* __restrict__ dZ, int lddz,
    float* __restrict__ dscale, float* __restrict__ dbias, int M, int N, int act) {
  pdl_prologue();
  __shared__ float sh_s[8][128 * V], sh_b[8][128 * V];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, wpb = blockDim.x >> 5;
  const float4* sc4 = reinterpret_cast<const float4*>(scale);
  const float4* bi4 = reinterpret_cast<const float4*>(bias);
  float gs[V][4], gb[V][4];
#pragma unroll
  for (int i = 0; i < V; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) gs[i][j] = gb[i][j] = 0.f;
  for (int64_t m = blockIdx.x * (int64_t)wpb + warp; m < M; m += (int64_t)gridDim.x * wpb) {
    const float mean = stats[2 * m], rstd = stats[2 * m + 1];
    const float4* z4 = reinterpret_cast<const float4*>(Z + (size_t)m * ldz);
    const float4* d4 = reinterpret_cast<const float4*>(dH + (size_t)m * lddh);
    float xh[V][4], dx[V][4], s1 = 0.f, s2 = 0.f;
#pragma unroll
    for (int i = 0; i < V; ++i) {
      const float4 z = z4[lane + 32 * i], dh = d4[lane + 32 * i];
      const float4 sc = sc4[lane + 32 * i];
      const float zz[4] = {z.x, z.y, z.z, z.w}, dd[4] = {dh.x, dh.y, dh.z, dh.w};
      const float ss[4] = {sc.x, sc.y, sc.z, sc.w};
      float bb[4] = {0.f, 0.f, 0.f, 0.f};
      if (act) {
        const float4 bi = bi4[lane + 32 * i];
        bb[0] = bi.x; bb[1] = bi.y; bb[2] = bi.z; bb[3] = bi.w;
      }
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float x = (zz[j] - mean) * rstd;
        float dy = dd[j];
        if (act) {
          const float y = x * ss[j] + bb[j];
          if (y <= 0.f) dy *= expf(y);
        }
        gs[i][j] += dy * x;
        gb[i][j] += dy;
        xh[i][j] = x;
        dx[i][j] = dy * ss[j];
        s1 += dx[i][j];
        s2 += dx[i][j] * x;
      }
    }
    const float m1 = warp_sum(s1) / N, m2 = warp_sum(s2) / N;
    float4* o4 = reinterpret_cast<float4*>(dZ + (size_t)m * lddz);
#pragma unroll
    for (int i = 0; i < V; ++i) {
      float4 o;
      o.x = rstd * (dx[i][0] - m1 - xh[i][0] * m2);
      o.y = rstd * (dx[i][1] - m1 - xh[i][1] * m2);
      o.z = rstd * (dx[i][2] - m1 - xh[i][2] * m2);
      o.w = rstd * (dx[i][3] - m1 - xh[i][3] * m2);
      o4[lane + 32 * i] = o;
    }
  }
  if (dscale == nullptr) return;
#pragma unroll
  for (int i = 0; i < V; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      sh_s[warp][4 * (lane + 32 * i) + j] = gs[i][j];
      sh_b[warp][4 * (lane + 32 * i) + j] = gb[i][j];
    }
  __syncthreads();
  for (int c = threadIdx.x; c < N; c += blockDim.x) {
    float a = 0.f, b = 0.f;
    for (int w = 0; w < wpb; ++w) {
      a += sh_s[w][c];
      b += sh_b[w][c];
    }
    atomicAdd(&dscale[c], a);
    atomicAdd(&dbias[c], b);
  }
}

static bool ln_vec_ok(const void* a, int lda, const void* b, int ldb, const void* c, const void* d,
                      int N) {
  auto al = [](const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; };
  return g_ln_vec && N % 128 == 0 && N <= 512 && (lda & 3) == 0 && (ldb & 3) == 0 && al(a) && al(b) &&
         al(c) && al(d);
}

}  // namespace dd

using namespace dd;

extern "C" int dd_ln_fwd(const float* Z, int ldz, const float* scale, const float* bias, float* H,
                         int ldh, float* stats, int M, int N, int act, dd_stream_t stream) {
  if (M <= 0) return 0;
  if (N > 12000) return fail("dd_ln_fwd", "row too wide for shared-memory staging");
  if (ln_vec_ok(Z, ldz, H, ldh, scale, bias, N)) {
    const int grid = grid_for(M, 8, 148 * 4);
#define DD_LNV(V) launch_k(ln_fwd_vec_kernel<V>, grid, 256, 0, S(stream), Z, ldz, scale, bias, H, ldh, stats, M, N, act)
    if (N == 128) DD_LNV(1);
    else if (N == 256) DD_LNV(2);
    else if (N == 384) DD_LNV(3);
    else DD_LNV(4);
#undef DD_LNV
    return launched("dd_ln_fwd");
  }
  if (N <= g_ln_warp_max) {
    const int grid = grid_for(M, 8, 148 * 16);
#define DD_LNF(PER) launch_k(ln_fwd_warp_kernel<PER>, grid, 256, 0, S(stream), Z, ldz, scale, bias, H, ldh, stats, M, N, act)
    if (N <= 32) DD_LNF(1);
    else if (N <= 64) DD_LNF(2);
    else if (N <= 128) DD_LNF(4);
    else if (N <= 256) DD_LNF(8);
    else DD_LNF(16);
#undef DD_LNF
    return launched("dd_ln_fwd");
  }
  const int block = N >= 2048 ? 512 : (N >= 256 ? 128 : 64);
  launch_k(ln_fwd_kernel, min(M, 148 * 32), block, N * sizeof(float), S(stream), 
      Z, ldz, scale, bias, H, ldh, stats, M, N, act);
  return launched("dd_ln_fwd");
}

extern "C" int dd_ln_bwd(const float* dH, int lddh, const float* Z, int ldz, const float* stats,
                         const float* scale, const float* bias, float* dZ, int lddz, float* dscale,
                         float* dbias, int M, int N, int act, dd_stream_t stream) {
  if (M <= 0) return 0;
  const int grid = min(M, 148 * 8);
#define DD_LNB(PER, BLK)                                                                      \
  launch_k(ln_bwd_kernel<PER, BLK>, grid, BLK, 0, S(stream), dH, lddh, Z, ldz, stats, scale, bias, dZ, \
                                                  lddz, dscale, dbias, M, N, act)
  if (ln_vec_ok(dH, lddh, Z, ldz, scale, bias, N) && (lddz & 3) == 0 &&
      (reinterpret_cast<uintptr_t>(dZ) & 15) == 0) {
    const int vgrid = grid_for(M, 8, 148 * 2);
#define DD_LNBV(V) launch_k(ln_bwd_vec_kernel<V>, vgrid, 256, 0, S(stream), dH, lddh, Z, ldz, stats, scale, bias, dZ, lddz, dscale, dbias, M, N, act)
    if (N == 128) DD_LNBV(1);
    else if (N == 256) DD_LNBV(2);
    else if (N == 384) DD_LNBV(3);
    else DD_LNBV(4);
#undef DD_LNBV
    return launched("dd_ln_bwd");
  }
  if (N <= g_ln_warp_max) {
    const int wgrid = grid_for(M, 8, 148 * 8);
#define DD_LNW(PER)                                                                              \
  launch_k(ln_bwd_warp_kernel<PER>, wgrid, 256, 0, S(stream), dH, lddh, Z, ldz, stats, scale, bias, \
           dZ, lddz, dscale, dbias, M, N, act)
    if (N <= 32) DD_LNW(1);
    else if (N <= 64) DD_LNW(2);
    else if (N <= 128) DD_LNW(4);
    else if (N <= 256) DD_LNW(8);
    else DD_LNW(16);
#undef DD_LNW
  }
  else if (N <= 128) DD_LNB(1, 128);
  else if (N <= 512) DD_LNB(4, 128);
  else if (N <= 2048) DD_LNB(8, 256);
  else if (N <= 8192) DD_LNB(8, 1024);
  else return fail("dd_ln_bwd", "N > 8192 unsupported");
#undef DD_LNB
  return launched("dd_ln_bwd");
}

extern "C" int dd_gru_fwd(const float* Y, int ldy, const float* prev, int ldp, float* out, int ldo,
                          int M, int D, dd_stream_t stream) {
  if (M <= 0) return 0;
  launch_k(gru_fwd_kernel, grid_for((int64_t)M * D, 256), 256, 0, S(stream), Y, ldy, prev, ldp, out,
                                                                       ldo, M, D);
  return launched("dd_gru_fwd");
}

extern "C" int dd_gru_bwd(const float* dOut, int lddo, const float* Y, int ldy, const float* prev,
                          int ldp, float* dY, int lddy, float* dPrev, int lddp, int M, int D,
                          int acc_prev, dd_stream_t stream) {
  if (M <= 0) return 0;
  launch_k(gru_bwd_kernel, grid_for((int64_t)M * D, 256), 256, 0, S(stream), 
      dOut, lddo, Y, ldy, prev, ldp, dY, lddy, dPrev, lddp, M, D, acc_prev);
  return launched("dd_gru_bwd");
}

extern "C" int dd_stats_fwd(const float* Z, int ldz, const float* G, int ldg, float* logit,
                            int ldl, float* stoch, int lds, int32_t* idx, int M, int S_, int C,
                            float mix, int mode, dd_stream_t stream) {
  if (M <= 0) return 0;
  if (C > 64) return fail("dd_stats_fwd", "classes > 64 unsupported");
  if (mode == 0 && G == nullptr) return fail("dd_stats_fwd", "mode 0 needs noise");
  if (mode != 2 && stoch == nullptr) return fail("dd_stats_fwd", "stoch output missing");
  const int64_t groups = (int64_t)M * S_;
  launch_k(stats_fwd_kernel, grid_for(groups, 8), 256, 0, S(stream), Z, ldz, G, ldg, logit, ldl, stoch,
                                                               lds, idx, M, S_, C, mix, mode);
  return launched("dd_stats_fwd");
}

extern "C" int dd_stats_bwd(const float* dStoch, int ldds, const float* dLogit, int lddl,
                            const float* logit, int ldl, float* dZ, int lddz, int M, int S_, int C,
                            float mix, dd_stream_t stream) {
  if (M <= 0) return 0;
  if (C > 64) return fail("dd_stats_bwd", "classes > 64 unsupported");
  const int64_t groups = (int64_t)M * S_;
  launch_k(stats_bwd_kernel, grid_for(groups, 8), 256, 0, S(stream), dStoch, ldds, dLogit, lddl, logit,
                                                               ldl, dZ, lddz, M, S_, C, mix);
  return launched("dd_stats_bwd");
}

extern "C" int dd_kl_fwd(const float* post, int ldp, const float* prior, int ldq, int M, int S_,
                         int C, float* kl_row, float* ent_post, float* ent_prior,
                         dd_stream_t stream) {
  if (M <= 0) return 0;
  if (C > 64) return fail("dd_kl_fwd", "classes > 64 unsupported");
  launch_k(kl_fwd_kernel, min(M, 148 * 16), 256, 0, S(stream), post, ldp, prior, ldq, M, S_, C, kl_row,
                                                         ent_post, ent_prior);
  return launched("dd_kl_fwd");
}

extern "C" int dd_kl_bwd(const float* post, int ldp, const float* prior, int ldq, int M, int S_,
                         int C, float bal, float coef, const float* scale, float* dpost, int lddp,
                         float* dprior, int lddq, dd_stream_t stream) {
  if (M <= 0) return 0;
  if (C > 64) return fail("dd_kl_bwd", "classes > 64 unsupported");
  const int64_t groups = (int64_t)M * S_;
  launch_k(kl_bwd_kernel, grid_for(groups, 8), 256, 0, S(stream), post, ldp, prior, ldq, M, S_, C, bal,
                                                            coef, scale, dpost, lddp, dprior, lddq);
  return launched("dd_kl_bwd");
}

extern "C" int dd_mask_fwd(const float* pd, int ldpd, const float* ps, int ldps, const float* act,
                           int lda, const float* first, const float* init_d, const float* init_s,
                           float* od, int ldod, float* os, int ldos, float* oa, int ldoa, int M,
                           int D, int SC, int A, dd_stream_t stream) {
  if (M <= 0) return 0;
  launch_k(mask_fwd_kernel, grid_for((int64_t)M * (D + SC + A), 256), 256, 0, S(stream), 
      pd, ldpd, ps, ldps, act, lda, first, init_d, init_s, od, ldod, os, ldos, oa, ldoa, M, D, SC,
      A);
  return launched("dd_mask_fwd");
}

extern "C" int dd_mask_bwd(const float* dd_, int lddd, const float* ds, int ldds,
                           const float* first, float* dpd, int lddpd, float* dps, int lddps,
                           float* dinit, int M, int D, int SC, dd_stream_t stream) {
  if (M <= 0) return 0;
  launch_k(mask_bwd_kernel, grid_for(D + SC, 128), 128, 0, S(stream), dd_, lddd, ds, ldds, first, dpd,
                                                                lddpd, dps, lddps, dinit, M, D, SC);
  return launched("dd_mask_bwd");
}

extern "C" int dd_set_ln_warp_max(int n) {
  const int prev = g_ln_warp_max;
  g_ln_warp_max = n < 0 ? 0 : (n > 512 ? 512 : n);
  return prev;
}

extern "C" int dd_set_ln_vec(int on) {
  const int prev = g_ln_vec;
  g_ln_vec = on ? 1 : 0;
  return prev;
}
