// Fused RSSM observe scan, forward (reference nets.py:66-76 `observe` unrolling
// `obs_step` :99-117 -> `img_step` :119-138 -> `_gru` :149-160 -> posterior).
//
// One persistent cooperative kernel runs all L steps.  The weights of the
// recurrent path stay in shared memory for the whole scan, column-sliced across
// the CTAs (weight-stationary):
//     img_in   W[(SC+A) x U]   -> U/G columns per CTA      (one-hot gather + A MACs)
//     gru_out  W[(D+U) x 3D]   -> the 3 gate columns of D/G deter units per CTA
//     obs_out  W[D x U] (deter part; the embed part is a batched GEMM done before)
//     obs_stats W[U x SC]      -> whole categorical groups per CTA
// Per step the M = B rows (B <= 16 sequences per GPU) flow through 5 phases
// separated by grid barriers; activations are exchanged through L2 (a few tens
// of KB), never through HBM-sized buffers, and LayerNorm / elu / GRU gates /
// unimix-softmax / Gumbel-argmax are fused into the phases.  Everything the
// backward pass needs is written once, time-major.
//
// Why this shape: with B = 16 rows the scan is pure latency; as
// ~12 dependent launches per step it cost ~85 us/step (5 us launch floor each);
// fused it is ~5 grid barriers + ~3 us of math per step.
#include <cooperative_groups.h>

#include "dd_common.cuh"

namespace dd {
namespace scan {

constexpr int MT = 16;          // max rows (sequences per GPU)
constexpr int NJ = 4;           // deter units per CTA (x3 gate columns)
constexpr int NU = 4;           // img_in / obs_out columns per CTA
constexpr int THREADS = 256;
constexpr int WARPS = THREADS / 32;

struct Smem {
  float* win;    // [(SC+A)][NU]
  float* wgru;   // [(D+U)][3*NJ]
  float* wobs;   // [D][NU]
  float* wst;    // [ngrp][U][CP]   CP = C + 4 (bank-conflict padding)
  float* x;      // [MT][XK]  staging of the current phase's input rows
  int* init_idx; // [S]      hot class of each group of the learned initial stoch
  int* idx0;     // [MT][S]  hot classes of the carried-in stoch (step 0 only)
};

__device__ __forceinline__ void grid_barrier(unsigned int* counter, unsigned int& epoch,
                                             unsigned int G) {
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence();
    atomicAdd(counter, 1u);
    const unsigned int target = (epoch + 1) * G;
    // back off between polls: 100+ CTAs hammering one L2 line slows every other
    // L2 access of the CTAs still working (measured: phases 3-4x slower without it)
    while (*reinterpret_cast<volatile unsigned int*>(counter) < target) {
      __nanosleep(32);
    }
    __threadfence();
  }
  epoch += 1;
  __syncthreads();
}

// rows r (2 per warp) x NC local columns, K split over lanes; W row stride ldw.
template <int NC>
__device__ __forceinline__ void rows_gemm(const float* __restrict__ X, int ldx, int K,
                                          const float* __restrict__ W, int ldw, int warp, int lane,
                                          float (&acc)[2][NC]) {
#pragma unroll
  for (int r = 0; r < 2; ++r)
#pragma unroll
    for (int c = 0; c < NC; ++c) acc[r][c] = 0.f;
  const float* x0 = X + (size_t)(2 * warp) * ldx;
  const float* x1 = x0 + ldx;
  for (int k = lane; k < K; k += 32) {
    const float a0 = x0[k], a1 = x1[k];
    const float* w = W + (size_t)k * ldw;
#pragma unroll
    for (int c4 = 0; c4 < NC / 4; ++c4) {
      const float4 v = *reinterpret_cast<const float4*>(w + 4 * c4);
      acc[0][4 * c4 + 0] = fmaf(a0, v.x, acc[0][4 * c4 + 0]);
      acc[0][4 * c4 + 1] = fmaf(a0, v.y, acc[0][4 * c4 + 1]);
      acc[0][4 * c4 + 2] = fmaf(a0, v.z, acc[0][4 * c4 + 2]);
      acc[0][4 * c4 + 3] = fmaf(a0, v.w, acc[0][4 * c4 + 3]);
      acc[1][4 * c4 + 0] = fmaf(a1, v.x, acc[1][4 * c4 + 0]);
      acc[1][4 * c4 + 1] = fmaf(a1, v.y, acc[1][4 * c4 + 1]);
      acc[1][4 * c4 + 2] = fmaf(a1, v.z, acc[1][4 * c4 + 2]);
      acc[1][4 * c4 + 3] = fmaf(a1, v.w, acc[1][4 * c4 + 3]);
    }
  }
#pragma unroll
  for (int r = 0; r < 2; ++r)
#pragma unroll
    for (int c = 0; c < NC; ++c) acc[r][c] = warp_sum(acc[r][c]);
}

// Each warp owns rows 2w and 2w+1 through every phase (staging, LayerNorm and the
// K-split GEMM only ever touch the warp's own rows), so staging needs no block
// barrier.  Loads come from L2 (ld.global.cg: other CTAs wrote them this step),
// as float4 when the row is 16-byte aligned.
template <int UN>
__device__ __forceinline__ void stage_my_rows(float* X, int ldx, const float* g, int ld, int B,
                                              int N, int warp, int lane) {
  const bool vec = ((ld & 3) == 0) && ((N & 3) == 0) && ((ldx & 3) == 0) &&
                   ((reinterpret_cast<uintptr_t>(g) & 15) == 0) &&
                   ((reinterpret_cast<uintptr_t>(X) & 15) == 0);
  const int b0 = 2 * warp, b1 = b0 + 1;
  float* x0 = X + (size_t)b0 * ldx;
  float* x1 = x0 + ldx;
  // rows past B read row 0 and are zeroed on the way into shared memory
  const float* s0 = g + (size_t)(b0 < B ? b0 : 0) * ld;
  const float* s1 = g + (size_t)(b1 < B ? b1 : 0) * ld;
  const float m0 = b0 < B ? 1.f : 0.f, m1 = b1 < B ? 1.f : 0.f;
  if (vec) {
    // both rows, UN float4 each: 2*UN independent L2 loads in flight per lane per round trip
    // (r01: with 4 in flight the 1536-wide rows of the BPTT kernel cost 6 round trips)
    for (int k0 = lane * 4; k0 < N; k0 += 128 * UN) {
      float4 a[UN], c[UN];
#pragma unroll
      for (int u = 0; u < UN; ++u) {
        const int k = k0 + u * 128;
        if (k < N) {
          a[u] = __ldcg(reinterpret_cast<const float4*>(s0 + k));
          c[u] = __ldcg(reinterpret_cast<const float4*>(s1 + k));
        }
      }
#pragma unroll
      for (int u = 0; u < UN; ++u) {
        const int k = k0 + u * 128;
        if (k < N) {
          *reinterpret_cast<float4*>(x0 + k) = make_float4(a[u].x * m0, a[u].y * m0, a[u].z * m0, a[u].w * m0);
          *reinterpret_cast<float4*>(x1 + k) = make_float4(c[u].x * m1, c[u].y * m1, c[u].z * m1, c[u].w * m1);
        }
      }
    }
  } else {
    for (int k = lane; k < N; k += 32) {
      x0[k] = __ldcg(s0 + k) * m0;
      x1[k] = __ldcg(s1 + k) * m1;
    }
  }
  __syncwarp();
}

// Two-pass LayerNorm statistics of row r of X (N values), computed by one warp.
__device__ __forceinline__ void row_stats(const float* x, int N, int lane, float& mean,
                                          float& rstd) {
  // 4 independent chains per pass: with 2 warps per scheduler a single dependent
  // LDS+FADD chain is pure latency
  float s[4] = {0.f, 0.f, 0.f, 0.f};
  int k = lane;
  for (; k + 96 < N; k += 128) {
    s[0] += x[k]; s[1] += x[k + 32]; s[2] += x[k + 64]; s[3] += x[k + 96];
  }
  for (; k < N; k += 32) s[0] += x[k];
  mean = warp_sum((s[0] + s[1]) + (s[2] + s[3])) / N;
  float q[4] = {0.f, 0.f, 0.f, 0.f};
  k = lane;
  for (; k + 96 < N; k += 128) {
    const float d0 = x[k] - mean, d1 = x[k + 32] - mean, d2 = x[k + 64] - mean, d3 = x[k + 96] - mean;
    q[0] = fmaf(d0, d0, q[0]); q[1] = fmaf(d1, d1, q[1]);
    q[2] = fmaf(d2, d2, q[2]); q[3] = fmaf(d3, d3, q[3]);
  }
  for (; k < N; k += 32) {
    const float d = x[k] - mean;
    q[0] = fmaf(d, d, q[0]);
  }
  rstd = rsqrtf(warp_sum((q[0] + q[1]) + (q[2] + q[3])) / N + DD_LN_EPS);
}

__global__ void __launch_bounds__(THREADS, 1) observe_fwd_kernel(DdObserveFwd p) {
  extern __shared__ __align__(16) float smem_f[];
  const int B = p.B, L = p.L, D = p.D, U = p.U, S = p.S, C = p.C, A = p.A;
  const int SC = S * C, G = gridDim.x, cta = blockIdx.x;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int CP = C + 4;
  const int ngrp = (S + G - 1) / G;                    // groups this grid assigns per CTA (max)
  const int XK = D + U;                                // widest staged row
  Smem sm;
  sm.win = smem_f;
  sm.wgru = sm.win + (size_t)(SC + A) * NU;
  sm.wobs = sm.wgru + (size_t)(D + U) * 3 * NJ;
  sm.wst = sm.wobs + (size_t)D * NU;
  sm.x = sm.wst + (size_t)ngrp * U * CP;
  sm.init_idx = reinterpret_cast<int*>(sm.x + (size_t)MT * XK);
  sm.idx0 = sm.init_idx + S;

  const int u0 = cta * NU;                             // this CTA's img_in / obs_out columns
  const int j0 = cta * NJ;                             // this CTA's deter units
  // ---- one-time: weight slices into shared memory --------------------------------
  for (int e = tid; e < (SC + A) * NU; e += THREADS) {
    const int k = e / NU, c = e % NU;
    sm.win[e] = (u0 + c < U) ? p.Win[(size_t)k * U + u0 + c] : 0.f;
  }
  for (int e = tid; e < (D + U) * 3 * NJ; e += THREADS) {
    const int k = e / (3 * NJ), lc = e % (3 * NJ);
    const int gate = lc / NJ, j = j0 + lc % NJ;
    sm.wgru[e] = (j < D) ? p.Wgru[(size_t)k * 3 * D + gate * D + j] : 0.f;
  }
  for (int e = tid; e < D * NU; e += THREADS) {
    const int k = e / NU, c = e % NU;
    sm.wobs[e] = (u0 + c < U) ? p.Wobs[(size_t)k * U + u0 + c] : 0.f;
  }
  for (int gi = 0; gi < ngrp; ++gi) {
    const int g = cta + gi * G;
    for (int e = tid; e < U * CP; e += THREADS) {
      const int k = e / CP, c = e % CP;
      sm.wst[(size_t)gi * U * CP + e] = (g < S && c < C) ? p.Wst[(size_t)k * SC + g * C + c] : 0.f;
    }
  }
  // hot classes of the initial / carried-in one-hot latents (found once, loads independent)
  for (int e = tid; e < S * (1 + MT); e += THREADS) {
    const int b = e / S - 1, sg = e % S;               // b == -1: the learned initial state
    const float* src = b < 0 ? p.init_stoch + sg * C : p.prev_stoch + (size_t)b * SC + sg * C;
    int hot = 0;
    if (b < B) {
      for (int c = C - 1; c >= 0; --c)
        if (src[c] > 0.5f) hot = c;
    }
    if (b < 0) sm.init_idx[sg] = hot; else sm.idx0[b * S + sg] = hot;
  }
  __syncthreads();

  unsigned int epoch = 0;
  const unsigned int uG = (unsigned int)G;
  // optional phase timing (profiling aid): slots [P1,B1,P2,B2,P3,B3,P4,B4,P5,B5]
  long long tacc[16] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0};
  long long tlast = clock64();
#define DD_TICK(slot)                                   \
  do {                                                  \
    const long long now_ = clock64();                   \
    tacc[slot] += now_ - tlast;                         \
    tlast = now_;                                       \
  } while (0)

  for (int t = 0; t < L; ++t) {
    const size_t row0 = (size_t)t * B;                 // first row of this step in [L*B, .] buffers
    const float* first = t == 0 ? p.first0 : p.first + row0;
    const float* pd = t == 0 ? p.prev_deter : p.feat + (row0 - B) * (D + SC);
    const int ldpd = t == 0 ? D : D + SC;
    const float* ps = t == 0 ? p.prev_stoch : p.feat + (row0 - B) * (D + SC) + D;
    const int ldps = t == 0 ? SC : D + SC;

    // ===== P1: masking + img_in (one-hot gather + action MACs) for columns u0.. =====
    // (a) every CTA writes its slice of the masked carry (saved for backward)
    for (int e = cta * THREADS + tid; e < B * (D + SC + A); e += G * THREADS) {
      const int b = e / (D + SC + A);
      int j = e % (D + SC + A);
      const float f = first[b], k1 = 1.f - f;
      if (j < D) {
        p.m_deter[(row0 + b) * D + j] = __ldcg(pd + (size_t)b * ldpd + j) * k1 + p.init_deter[j] * f;
      } else if ((j -= D) < SC) {
        p.m_stoch[(row0 + b) * SC + j] = __ldcg(ps + (size_t)b * ldps + j) * k1 + p.init_stoch[j] * f;
      } else {
        j -= SC;
        p.m_act[(row0 + b) * A + j] = p.action[(row0 + b) * A + j] * k1;
      }
    }
    // (b) z_in[b, u0+c] = sum_k m_stoch[b,k] Win[k,c] + sum_a m_act[b,a] Win[SC+a,c]; the masked
    //     stoch is one-hot per group when first is 0/1, but stay general (dense over SC) only
    //     where needed: warp w handles rows 2w, 2w+1; lanes split groups.
    {
      float acc[2][NU];
#pragma unroll
      for (int r = 0; r < 2; ++r)
#pragma unroll
        for (int c = 0; c < NU; ++c) acc[r][c] = 0.f;
#pragma unroll
      for (int r = 0; r < 2; ++r) {
        const int b = 2 * warp + r;
        if (b < B) {
          const float f = first[b];
          for (int sg = lane; sg < S; sg += 32) {
            // hot class of group sg of the masked (one-hot) stoch
            int hot;
            if (f >= 0.5f) hot = sm.init_idx[sg];
            else if (t == 0) hot = sm.idx0[b * S + sg];
            else hot = __ldcg(p.post_idx + (row0 - B + b) * S + sg);
            const float4 w = *reinterpret_cast<const float4*>(sm.win + (size_t)(sg * C + hot) * NU);
            acc[r][0] += w.x; acc[r][1] += w.y; acc[r][2] += w.z; acc[r][3] += w.w;
          }
          for (int a = lane; a < A; a += 32) {
            const float av = p.action[(row0 + b) * A + a] * (1.f - f);
            const float4 w = *reinterpret_cast<const float4*>(sm.win + (size_t)(SC + a) * NU);
            acc[r][0] = fmaf(av, w.x, acc[r][0]); acc[r][1] = fmaf(av, w.y, acc[r][1]);
            acc[r][2] = fmaf(av, w.z, acc[r][2]); acc[r][3] = fmaf(av, w.w, acc[r][3]);
          }
        }
      }
#pragma unroll
      for (int r = 0; r < 2; ++r)
#pragma unroll
        for (int c = 0; c < NU; ++c) {
          const float v = warp_sum(acc[r][c]);
          const int b = 2 * warp + r;
          if (lane == 0 && b < B && u0 + c < U) p.z_in[(row0 + b) * U + u0 + c] = v;
        }
    }
    DD_TICK(0);
    grid_barrier(p.barrier, epoch, uG);
    DD_TICK(1);

    // ===== P2: x = elu(LN(z_in)); z_gru[:, my gate cols] = [m_deter | x] @ Wgru slice =====
    stage_my_rows<2>(sm.x + D, XK, p.z_in + row0 * U, U, B, U, warp, lane);
    stage_my_rows<2>(sm.x, XK, p.m_deter + row0 * D, D, B, D, warp, lane);
    DD_TICK(12);
    {
#pragma unroll
      for (int r = 0; r < 2; ++r) {
        const int b = 2 * warp + r;
        float* xr = sm.x + (size_t)b * XK + D;
        float mean, rstd;
        row_stats(xr, U, lane, mean, rstd);
        for (int k = lane; k < U; k += 32) {
          float y = (xr[k] - mean) * rstd * p.ln_in_s[k] + p.ln_in_b[k];
          y = y > 0.f ? y : expm1f(y);
          xr[k] = y;
          if (cta == warp % G && b < B) p.x_in[(row0 + b) * U + k] = y;
        }
        if (cta == warp % G && lane == 0 && b < B) {
          p.st_in[(row0 + b) * 2] = mean;
          p.st_in[(row0 + b) * 2 + 1] = rstd;
        }
      }
    }
    __syncwarp();
    float zg[2][3 * NJ];                                // this CTA's gate pre-activations (kept)
    rows_gemm<3 * NJ>(sm.x, XK, D + U, sm.wgru, 3 * NJ, warp, lane, zg);
    {
      // partial LayerNorm sums over this CTA's 3*NJ columns, and the saved z_gru
#pragma unroll
      for (int r = 0; r < 2; ++r) {
        const int b = 2 * warp + r;
        // local (count, sum, M2) over this CTA's valid gate columns
        float s1 = 0.f;
        int cnt = 0;
#pragma unroll
        for (int lc = 0; lc < 3 * NJ; ++lc)
          if (j0 + lc % NJ < D) { s1 += zg[r][lc]; ++cnt; }
        const float lm = cnt > 0 ? s1 / cnt : 0.f;
        float m2 = 0.f;
#pragma unroll
        for (int lc = 0; lc < 3 * NJ; ++lc)
          if (j0 + lc % NJ < D) { const float d = zg[r][lc] - lm; m2 += d * d; }
        if (lane == 0 && b < B) {
          p.partial[((size_t)b * 148 + cta) * 2] = s1;
          p.partial[((size_t)b * 148 + cta) * 2 + 1] = m2;
#pragma unroll
          for (int lc = 0; lc < 3 * NJ; ++lc) {
            const int gate = lc / NJ, j = j0 + lc % NJ;
            if (j < D) p.z_gru[(row0 + b) * 3 * D + gate * D + j] = zg[r][lc];
          }
        }
      }
    }
    DD_TICK(2);
    grid_barrier(p.barrier, epoch, uG);
    DD_TICK(3);

    // ===== P3: joint LayerNorm over 3D + GRU gates for my deter units ================
    {
      // issue every L2 load of both rows before the first use
      float ps1[2][(148 + 31) / 32], ps2[2][(148 + 31) / 32], prevv[2];
#pragma unroll
      for (int r = 0; r < 2; ++r) {
        const int b = 2 * warp + r;
#pragma unroll
        for (int i = 0; i < (148 + 31) / 32; ++i) {
          const int c = lane + 32 * i;
          // [row][cta][2]: a warp reads 256 contiguous bytes (the first layout, [cta][row],
          // made every lane hit a different line that all 128 CTAs wanted: 8 us per step)
          const float2 v = c < G ? __ldcg(reinterpret_cast<const float2*>(p.partial) + (size_t)b * 148 + c)
                                 : make_float2(0.f, 0.f);
          ps1[r][i] = v.x;
          ps2[r][i] = v.y;
        }
        prevv[r] = (b < B && lane < NJ && j0 + lane < D)
                       ? __ldcg(p.m_deter + (row0 + b) * D + j0 + lane) : 0.f;
      }
      if (ps1[0][0] + ps2[1][0] + prevv[0] + prevv[1] == 123.456f) tlast += 1;   // force the loads to land
      DD_TICK(10);
#pragma unroll
      for (int r = 0; r < 2; ++r) {
        const int b = 2 * warp + r;
        // merge the per-CTA (sum, M2) partials: mean first, then M2 (Chan et al.), all fp32
        float s1 = 0.f;
#pragma unroll
        for (int i = 0; i < (148 + 31) / 32; ++i) s1 += ps1[r][i];
        const float mean = warp_sum(s1) / (3.f * D);
        float m2 = 0.f;
#pragma unroll
        for (int i = 0; i < (148 + 31) / 32; ++i) {
          const int c = lane + 32 * i;
          const int nv = 3 * max(0, min(NJ, D - c * NJ));          // valid columns of CTA c
          if (c < G && nv > 0) {
            const float lm = ps1[r][i] / nv;
            m2 += ps2[r][i] + nv * (lm - mean) * (lm - mean);
          }
        }
        const float rstd = rsqrtf(warp_sum(m2) / (3.f * D) + DD_LN_EPS);
        if (b < B && lane < NJ && j0 + lane < D) {
          const int j = j0 + lane;
          float y[3];
#pragma unroll
          for (int gate = 0; gate < 3; ++gate) {
            // register arrays are indexed with compile-time constants only
            float z = 0.f;
#pragma unroll
            for (int q = 0; q < NJ; ++q)
              if (lane == q) z = zg[r][gate * NJ + q];
            y[gate] = (z - mean) * rstd * p.ln_gru_s[gate * D + j] + p.ln_gru_b[gate * D + j];
            p.y_gru[(row0 + b) * 3 * D + gate * D + j] = y[gate];
          }
          const float rg = sigmoidf_(y[0]);
          const float cg = tanhf(rg * y[1]);
          const float ug = sigmoidf_(y[2] - 1.f);
          const float prev = prevv[r];
          p.feat[(row0 + b) * (D + SC) + j] = ug * cg + (1.f - ug) * prev;
        }
        if (cta == 0 && lane == 0 && b < B) {
          p.st_gru[(row0 + b) * 2] = mean;
          p.st_gru[(row0 + b) * 2 + 1] = rstd;
        }
      }
    }
    DD_TICK(4);
    grid_barrier(p.barrier, epoch, uG);
    DD_TICK(5);

    // ===== P4: z_obs[:, u0..] += deter_t @ Wobs[:D] slice  (embed part already there) ====
    stage_my_rows<2>(sm.x, XK, p.feat + row0 * (D + SC), D + SC, B, D, warp, lane);
    {
      float zo[2][NU];
      rows_gemm<NU>(sm.x, XK, D, sm.wobs, NU, warp, lane, zo);
#pragma unroll
      for (int r = 0; r < 2; ++r) {
        const int b = 2 * warp + r;
        if (lane == 0 && b < B) {
#pragma unroll
          for (int c = 0; c < NU; ++c)
            if (u0 + c < U) p.z_obs[(row0 + b) * U + u0 + c] += zo[r][c];
        }
      }
    }
    DD_TICK(6);
    grid_barrier(p.barrier, epoch, uG);
    DD_TICK(7);

    // ===== P5: h = elu(LN(z_obs)); logits of my groups; unimix; Gumbel-argmax ==========
    if (cta < S) {
      stage_my_rows<2>(sm.x, XK, p.z_obs + row0 * U, U, B, U, warp, lane);
#pragma unroll
      for (int r = 0; r < 2; ++r) {
        const int b = 2 * warp + r;
        float* xr = sm.x + (size_t)b * XK;
        float mean, rstd;
        row_stats(xr, U, lane, mean, rstd);
        for (int k = lane; k < U; k += 32) {
          float y = (xr[k] - mean) * rstd * p.ln_obs_s[k] + p.ln_obs_b[k];
          y = y > 0.f ? y : expm1f(y);
          xr[k] = y;
          if (cta == warp % min(S, G) && b < B) p.h_obs[(row0 + b) * U + k] = y;
        }
        if (cta == warp % min(S, G) && lane == 0 && b < B) {
          p.st_obs[(row0 + b) * 2] = mean;
          p.st_obs[(row0 + b) * 2 + 1] = rstd;
        }
      }
      __syncwarp();
      for (int gi = 0; gi < ngrp; ++gi) {
        const int g = cta + gi * G;
        if (g >= S) break;
        const float* wg = sm.wst + (size_t)gi * U * CP;
        // lanes = classes (C <= 64: two per lane), warps = row pairs; K serial over U with
        // 4 independent FMA chains per lane (2 rows x 2 k-phases)
        float za[2][2] = {{0.f, 0.f}, {0.f, 0.f}}, zb[2][2] = {{0.f, 0.f}, {0.f, 0.f}};
        if (C == 32) {
          // K split over lanes, 2 rows x 32 classes per lane, then lane c keeps class c
          float acc[2][32];
          rows_gemm<32>(sm.x, XK, U, wg, CP, warp, lane, acc);
#pragma unroll
          for (int c = 0; c < 32; ++c)
            if (lane == c) { za[0][0] = acc[0][c]; za[1][0] = acc[1][c]; }
        } else {
          const float* xr0 = sm.x + (size_t)(2 * warp) * XK;
          const float* xr1 = xr0 + XK;
          const int cc0 = lane < C ? lane : 0, cc1 = lane + 32 < C ? lane + 32 : 0;
          for (int k = 0; k + 1 < U; k += 2) {
            const float w00 = wg[(size_t)k * CP + cc0], w01 = wg[(size_t)(k + 1) * CP + cc0];
            za[0][0] = fmaf(xr0[k], w00, za[0][0]); za[0][1] = fmaf(xr0[k + 1], w01, za[0][1]);
            za[1][0] = fmaf(xr1[k], w00, za[1][0]); za[1][1] = fmaf(xr1[k + 1], w01, za[1][1]);
            if (C > 32) {
              const float w10 = wg[(size_t)k * CP + cc1], w11 = wg[(size_t)(k + 1) * CP + cc1];
              zb[0][0] = fmaf(xr0[k], w10, zb[0][0]); zb[0][1] = fmaf(xr0[k + 1], w11, zb[0][1]);
              zb[1][0] = fmaf(xr1[k], w10, zb[1][0]); zb[1][1] = fmaf(xr1[k + 1], w11, zb[1][1]);
            }
          }
          if (U & 1) {
            const int k = U - 1;
            za[0][0] = fmaf(xr0[k], wg[(size_t)k * CP + cc0], za[0][0]);
            za[1][0] = fmaf(xr1[k], wg[(size_t)k * CP + cc0], za[1][0]);
            if (C > 32) {
              zb[0][0] = fmaf(xr0[k], wg[(size_t)k * CP + cc1], zb[0][0]);
              zb[1][0] = fmaf(xr1[k], wg[(size_t)k * CP + cc1], zb[1][0]);
            }
          }
        }
#pragma unroll
        for (int r = 0; r < 2; ++r) {
          const int b = 2 * warp + r;
          const int c0 = lane, c1 = lane + 32;
          float z0 = za[r][0] + za[r][1], z1 = zb[r][0] + zb[r][1];
          z0 = c0 < C ? z0 + p.bst[g * C + c0] : -INFINITY;
          z1 = c1 < C ? z1 + p.bst[g * C + c1] : -INFINITY;
          const float mx = warp_max(fmaxf(z0, z1));
          const float e0 = c0 < C ? expf(z0 - mx) : 0.f;
          const float e1 = c1 < C ? expf(z1 - mx) : 0.f;
          const float inv = 1.f / warp_sum(e0 + e1);
          const float un = p.mix / C;
          const float l0 = logf((1.f - p.mix) * e0 * inv + un);
          const float l1 = logf((1.f - p.mix) * e1 * inv + un);
          if (b < B) {
            const float* gn = p.g_obs + (row0 + b) * SC + g * C;
            float bv = c0 < C ? l0 + gn[c0] : -INFINITY;
            int bi = c0 < C ? c0 : 0x7fffffff;
            if (c1 < C) {
              const float k1 = l1 + gn[c1];
              if (k1 > bv) { bv = k1; bi = c1; }
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
              const float ov = __shfl_xor_sync(0xffffffffu, bv, o);
              const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
              if (ov > bv || (ov == bv && oi < bi)) { bv = ov; bi = oi; }
            }
            float* lg = p.post_logit + (row0 + b) * SC + g * C;
            float* st = p.feat + (row0 + b) * (D + SC) + D + g * C;
            if (c0 < C) { lg[c0] = l0; st[c0] = (c0 == bi) ? 1.f : 0.f; }
            if (c1 < C) { lg[c1] = l1; st[c1] = (c1 == bi) ? 1.f : 0.f; }
            if (lane == 0) p.post_idx[(row0 + b) * S + g] = bi;
          }
        }
      }
    }
    DD_TICK(8);
    grid_barrier(p.barrier, epoch, uG);
    DD_TICK(9);
  }
#undef DD_TICK
  if (cta == 0 && tid == 0 && p.timing != nullptr)
    for (int i = 0; i < 16; ++i) p.timing[i] = tacc[i];
}

// ---------------------------------------------------------------------------
// Backward of the observe scan (BPTT), fused the same way: one persistent
// cooperative kernel walks t = L-1 .. 0.  Weight-stationary TRANSPOSED slices (the
// input-gradient GEMMs are dZ @ W^T, so a CTA that owns output columns owns rows
// of W): obs_stats rows of its NU units, obs_out rows of its NJ deter units, the
// gru_out rows of its NJ deter + NU x units, img_in rows of its stoch columns.
// Per step (reference math = engine.Engine.wm_backward's per-step launches:
// stats_bwd -> GEMM -> ln_bwd -> GEMM -> gru_bwd -> ln_bwd(joint 3D) -> 2 GEMMs ->
// ln_bwd -> GEMM -> mask_bwd) 6 phases / 6 grid barriers:
//   P1a straight-through + unimix backward of the posterior draw for my groups
//   P1b dh = dzs @ Wst^T
//   P2 LN backward (obs_out), d deter += dz_obs @ Wobs[:D]^T, GRU gate backward for
//      my deter units, partial sums of the joint 3D LayerNorm backward
//   P3 joint LN backward -> dz_gru for my gate columns
//   P4 [d prev_deter | dx] = dz_gru @ Wgru^T ; is_first masking of the deter carry
//   P5 LN backward (img_in), d prev_stoch = dz_in @ Win[:SC]^T ; masking
// dzs / dz_obs / dz_gru / dz_in and the pre-LN gradients dh / dy / dx are written
// time-major for the batched weight-gradient GEMMs and LN parameter gradients that
// follow the kernel.
// ---------------------------------------------------------------------------
constexpr int NSW = 4;          // stoch columns per img_in chunk

// LayerNorm(+elu) backward of rows 2w, 2w+1, one warp.  X[b][0:N] holds dH on entry and dZ on
// exit; X[b][N:2N] holds the saved pre-norm row Z (both staged in shared memory).
__device__ __forceinline__ void ln_bwd_rows(float* X, int ldx, int N, const float* st,
                                            const float* __restrict__ scale,
                                            const float* __restrict__ bias, int act, float* out,
                                            int B, int warp, int lane) {
  // both rows of the warp advance together: two independent dependency chains per lane
  const int b0 = 2 * warp, b1 = b0 + 1;
  const bool a0 = b0 < B, a1 = b1 < B;
  if (!a0) return;                                       // b1 >= B as well
  float* x0 = X + (size_t)b0 * ldx;
  float* x1 = x0 + ldx;
  const float* z0 = x0 + N;
  const float* z1 = x1 + N;
  const float mean0 = st[2 * b0], rstd0 = st[2 * b0 + 1];
  const float mean1 = a1 ? st[2 * b1] : 0.f, rstd1 = a1 ? st[2 * b1 + 1] : 0.f;
  float s10 = 0.f, s20 = 0.f, s11 = 0.f, s21 = 0.f;
#pragma unroll 4
  for (int k = lane; k < N; k += 32) {
    const float sc = scale[k], bi = bias[k];
    const float xh0 = (z0[k] - mean0) * rstd0, xh1 = (z1[k] - mean1) * rstd1;
    float dy0 = x0[k], dy1 = x1[k];
    if (act) {
      const float y0 = xh0 * sc + bi, y1 = xh1 * sc + bi;
      if (y0 <= 0.f) dy0 *= expf(y0);
      if (y1 <= 0.f) dy1 *= expf(y1);
    }
    const float d0 = dy0 * sc, d1 = dy1 * sc;
    x0[k] = d0;
    x1[k] = d1;
    s10 += d0; s20 += d0 * xh0;
    s11 += d1; s21 += d1 * xh1;
  }
  const float m10 = warp_sum(s10) / N, m20 = warp_sum(s20) / N;
  const float m11 = warp_sum(s11) / N, m21 = warp_sum(s21) / N;
#pragma unroll 4
  for (int k = lane; k < N; k += 32) {
    const float xh0 = (z0[k] - mean0) * rstd0, xh1 = (z1[k] - mean1) * rstd1;
    const float dz0 = rstd0 * (x0[k] - m10 - xh0 * m20);
    const float dz1 = a1 ? rstd1 * (x1[k] - m11 - xh1 * m21) : 0.f;
    x0[k] = dz0;
    x1[k] = dz1;
    if (out != nullptr) {
      out[(size_t)b0 * N + k] = dz0;
      if (a1) out[(size_t)b1 * N + k] = dz1;
    }
  }
  __syncwarp();
}

__global__ void __launch_bounds__(THREADS, 1) observe_bwd_kernel(DdObserveBwd p) {
  extern __shared__ __align__(16) float smem_f[];
  const int B = p.B, L = p.L, D = p.D, U = p.U, S = p.S, C = p.C;
  const int SC = S * C, F = D + SC, G = gridDim.x, cta = blockIdx.x;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int D3 = 3 * D;
  const int nsc = p.nsc;                               // img_in chunks (of NSW columns) per CTA
  const int XK = p.xk;                                 // staged row width (multiple of 4)
  float* wst = smem_f;                                 // [SC][NU]
  float* wobs = wst + (size_t)SC * NU;                 // [U][NJ]
  float* wgru = wobs + (size_t)U * NJ;                 // [3D][NJ + NU]
  float* win = wgru + (size_t)D3 * (NJ + NU);          // [nsc][U][NSW]
  float* xs = win + (size_t)nsc * U * NSW;             // [MT][XK]
  const int u0 = cta * NU, j0 = cta * NJ, n0 = cta * nsc * NSW;
  const bool owner = cta == warp % G;                  // this CTA stores the full rows of warp `warp`

  for (int e = tid; e < SC * NU; e += THREADS) {
    const int k = e / NU, c = e % NU;
    wst[e] = (u0 + c < U) ? p.Wst[(size_t)(u0 + c) * SC + k] : 0.f;
  }
  for (int e = tid; e < U * NJ; e += THREADS) {
    const int k = e / NJ, c = e % NJ;
    wobs[e] = (j0 + c < D) ? p.Wobs[(size_t)(j0 + c) * U + k] : 0.f;
  }
  for (int e = tid; e < D3 * (NJ + NU); e += THREADS) {
    const int k = e / (NJ + NU), c = e % (NJ + NU);
    float v = 0.f;
    if (c < NJ) { if (j0 + c < D) v = p.Wgru[(size_t)(j0 + c) * D3 + k]; }
    else if (u0 + c - NJ < U) v = p.Wgru[(size_t)(D + u0 + c - NJ) * D3 + k];
    wgru[e] = v;
  }
  for (int e = tid; e < nsc * U * NSW; e += THREADS) {
    const int ch = e / (U * NSW), k = (e / NSW) % U, c = e % NSW;
    const int n = n0 + ch * NSW + c;
    win[e] = (n < SC) ? p.Win[(size_t)n * U + k] : 0.f;
  }
  __syncthreads();

  unsigned int epoch = 0;
  const unsigned int uG = (unsigned int)G;
  const float un = p.mix / C, om = 1.f - p.mix;
  // optional phase timing (profiling aid): slots [P1a,B,P1b,B,P2,B,P3,B,P4,B,P5,B] of CTA 0
  long long tacc[12] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0};
  long long tlast = clock64();
  int tslot = 0;
#define DD_TICKB()                                      \
  do {                                                  \
    const long long now_ = clock64();                   \
    tacc[tslot] += now_ - tlast;                        \
    tlast = now_;                                       \
    tslot = (tslot + 1) % 12;                           \
  } while (0)

  for (int t = L - 1; t >= 0; --t) {
    const size_t row0 = (size_t)t * B;
    const float* first = t == 0 ? p.first0 : p.first + row0;

    // ===== P1a: dzs = d(raw posterior logits) for my categorical groups ===================
    // (straight-through + unimix backward; lanes = classes, warps = row pairs)
    for (int g = cta; g < S; g += G) {
      const int c0 = g * C + lane, c1 = c0 + 32;
      const bool v0 = lane < C, v1 = lane + 32 < C;
      float lg[2][2], dlg[2][2], dst[2][2];
#pragma unroll
      for (int r = 0; r < 2; ++r) {                      // every load of both rows up front
        const int b = 2 * warp + r;
        const size_t rr = row0 + (b < B ? b : 0);
        lg[r][0] = v0 ? p.post_logit[rr * SC + c0] : 0.f;
        lg[r][1] = v1 ? p.post_logit[rr * SC + c1] : 0.f;
        dlg[r][0] = v0 ? p.dpost_logit[rr * SC + c0] : 0.f;
        dlg[r][1] = v1 ? p.dpost_logit[rr * SC + c1] : 0.f;
        dst[r][0] = v0 ? __ldcg(p.dfeat + rr * F + D + c0) : 0.f;
        dst[r][1] = v1 ? __ldcg(p.dfeat + rr * F + D + c1) : 0.f;
      }
#pragma unroll
      for (int r = 0; r < 2; ++r) {
        const int b = 2 * warp + r;
        const float q0 = v0 ? expf(lg[r][0]) : 0.f, q1 = v1 ? expf(lg[r][1]) : 0.f;
        const float dot = warp_sum(q0 * dst[r][0] + q1 * dst[r][1]);
        const float dl0 = q0 * (dst[r][0] - dot) + dlg[r][0];
        const float dl1 = q1 * (dst[r][1] - dot) + dlg[r][1];
        const float p0 = v0 ? (q0 - un) / om : 0.f, p1 = v1 ? (q1 - un) / om : 0.f;
        const float dp0 = v0 ? om * dl0 / q0 : 0.f, dp1 = v1 ? om * dl1 / q1 : 0.f;
        const float dot2 = warp_sum(p0 * dp0 + p1 * dp1);
        if (b < B) {
          if (v0) p.dzs[(row0 + b) * SC + c0] = p0 * (dp0 - dot2);
          if (v1) p.dzs[(row0 + b) * SC + c1] = p1 * (dp1 - dot2);
        }
      }
    }
    DD_TICKB();
    grid_barrier(p.barrier, epoch, uG);
    DD_TICKB();

    // ===== P1b: dh[:, my units] = dzs @ Wst^T ============================================
    stage_my_rows<4>(xs, XK, p.dzs + row0 * SC, SC, B, SC, warp, lane);
    {
      float acc[2][NU];
      rows_gemm<NU>(xs, XK, SC, wst, NU, warp, lane, acc);
#pragma unroll
      for (int r = 0; r < 2; ++r) {
        const int b = 2 * warp + r;
        if (lane == 0 && b < B) {
#pragma unroll
          for (int c = 0; c < NU; ++c)
            if (u0 + c < U) p.dh[(row0 + b) * U + u0 + c] = acc[r][c];
        }
      }
    }
    DD_TICKB();
    grid_barrier(p.barrier, epoch, uG);
    DD_TICKB();

    // ===== P2: LN backward (obs_out) -> deter gradient -> GRU gates of my units ==========
    stage_my_rows<4>(xs, XK, p.dh + row0 * U, U, B, U, warp, lane);
    stage_my_rows<4>(xs + U, XK, p.z_obs + row0 * U, U, B, U, warp, lane);
    ln_bwd_rows(xs, XK, U, p.st_obs + row0 * 2, p.ln_obs_s, p.ln_obs_b, 1,
                owner ? p.dz_obs + row0 * U : nullptr, B, warp, lane);
    float ddir[2] = {0.f, 0.f};                        // direct path g*(1-u) of my deter unit
    float dxg[2][3], xhg[2][3], rstd_g[2] = {0.f, 0.f};
    {
      float acc[2][NJ];
      rows_gemm<NJ>(xs, XK, U, wobs, NJ, warp, lane, acc);
#pragma unroll
      for (int r = 0; r < 2; ++r) {
        const int b = 2 * warp + r;
        float s1 = 0.f, s2 = 0.f;
#pragma unroll
        for (int gate = 0; gate < 3; ++gate) dxg[r][gate] = xhg[r][gate] = 0.f;
        if (b < B && lane < NJ && j0 + lane < D) {
          const int j = j0 + lane;
          float gsum = 0.f;
#pragma unroll
          for (int q = 0; q < NJ; ++q)
            if (lane == q) gsum = acc[r][q];
          const float g = gsum + __ldcg(p.dfeat + (row0 + b) * F + j);
          const float* y = p.y_gru + (row0 + b) * D3;
          const float yr = y[j], yc = y[D + j], yu = y[2 * D + j];
          const float prev = p.m_deter[(row0 + b) * D + j];
          const float rg = sigmoidf_(yr), cg = tanhf(rg * yc), ug = sigmoidf_(yu - 1.f);
          const float dpre = g * ug * (1.f - cg * cg);
          float dy[3];
          dy[0] = dpre * yc * rg * (1.f - rg);
          dy[1] = dpre * rg;
          dy[2] = g * (cg - prev) * ug * (1.f - ug);
          ddir[r] = g * (1.f - ug);
          const float mean = p.st_gru[(row0 + b) * 2];
          rstd_g[r] = p.st_gru[(row0 + b) * 2 + 1];
#pragma unroll
          for (int gate = 0; gate < 3; ++gate) {
            const int col = gate * D + j;
            p.dy_gru[(row0 + b) * D3 + col] = dy[gate];
            xhg[r][gate] = (p.z_gru[(row0 + b) * D3 + col] - mean) * rstd_g[r];
            dxg[r][gate] = dy[gate] * p.ln_gru_s[col];
            s1 += dxg[r][gate];
            s2 += dxg[r][gate] * xhg[r][gate];
          }
        }
        s1 = warp_sum(s1);
        s2 = warp_sum(s2);
        if (lane == 0 && b < B) {
          p.partial[((size_t)b * 148 + cta) * 2] = s1;
          p.partial[((size_t)b * 148 + cta) * 2 + 1] = s2;
        }
      }
    }
    DD_TICKB();
    grid_barrier(p.barrier, epoch, uG);
    DD_TICKB();

    // ===== P3: joint LayerNorm backward over 3D for my gate columns =======================
#pragma unroll
    for (int r = 0; r < 2; ++r) {
      const int b = 2 * warp + r;
      float s1 = 0.f, s2 = 0.f;
#pragma unroll
      for (int i = 0; i < (148 + 31) / 32; ++i) {
        const int c = lane + 32 * i;
        if (c < G && b < B) {
          const float2 v = __ldcg(reinterpret_cast<const float2*>(p.partial) + (size_t)b * 148 + c);
          s1 += v.x;
          s2 += v.y;
        }
      }
      const float m1 = warp_sum(s1) / D3, m2 = warp_sum(s2) / D3;
      if (b < B && lane < NJ && j0 + lane < D) {
#pragma unroll
        for (int gate = 0; gate < 3; ++gate)
          p.dz_gru[(row0 + b) * D3 + gate * D + j0 + lane] =
              rstd_g[r] * (dxg[r][gate] - m1 - xhg[r][gate] * m2);
      }
    }
    DD_TICKB();
    grid_barrier(p.barrier, epoch, uG);
    DD_TICKB();

    // ===== P4: [d prev_deter | dx] = dz_gru @ Wgru^T ; deter carry masking ================
    stage_my_rows<4>(xs, XK, p.dz_gru + row0 * D3, D3, B, D3, warp, lane);
    {
      float acc[2][NJ + NU];
      rows_gemm<NJ + NU>(xs, XK, D3, wgru, NJ + NU, warp, lane, acc);
#pragma unroll
      for (int r = 0; r < 2; ++r) {
        const int b = 2 * warp + r;
        float mine = 0.f;
#pragma unroll
        for (int q = 0; q < NJ + NU; ++q)
          if (lane == q) mine = acc[r][q];
        if (b < B && lane < NJ && j0 + lane < D) {
          const int j = j0 + lane;
          const float dd = ddir[r] + mine, f = first[b];
          if (f != 0.f) atomicAdd(p.d_initdeter + j, dd * f);
          if (t > 0) {
            float* q = p.dfeat + (row0 - B + b) * F + j;
            *q = __ldcg(q) + dd * (1.f - f);
          }
        } else if (b < B && lane >= NJ && lane < NJ + NU && u0 + lane - NJ < U) {
          p.dx[(row0 + b) * U + u0 + lane - NJ] = mine;
        }
      }
    }
    DD_TICKB();
    grid_barrier(p.barrier, epoch, uG);
    DD_TICKB();

    // ===== P5: LN backward (img_in) ; d prev_stoch = dz_in @ Win[:SC]^T ; masking ==========
    stage_my_rows<4>(xs, XK, p.dx + row0 * U, U, B, U, warp, lane);
    stage_my_rows<4>(xs + U, XK, p.z_in + row0 * U, U, B, U, warp, lane);
    ln_bwd_rows(xs, XK, U, p.st_in + row0 * 2, p.ln_in_s, p.ln_in_b, 1,
                owner ? p.dz_in + row0 * U : nullptr, B, warp, lane);
    if (t > 0) {
      for (int ch = 0; ch < nsc; ++ch) {
        float acc[2][NSW];
        rows_gemm<NSW>(xs, XK, U, win + (size_t)ch * U * NSW, NSW, warp, lane, acc);
#pragma unroll
        for (int r = 0; r < 2; ++r) {
          const int b = 2 * warp + r;
          float mine = 0.f;
#pragma unroll
          for (int q = 0; q < NSW; ++q)
            if (lane == q) mine = acc[r][q];
          const int n = n0 + ch * NSW + lane;
          if (b < B && lane < NSW && n < SC) {
            float* q = p.dfeat + (row0 - B + b) * F + D + n;
            *q = __ldcg(q) + mine * (1.f - first[b]);
          }
        }
      }
      DD_TICKB();
      grid_barrier(p.barrier, epoch, uG);
      DD_TICKB();
    }
  }
#undef DD_TICKB
  if (cta == 0 && tid == 0 && p.timing != nullptr)
    for (int i = 0; i < 12; ++i) p.timing[i] = tacc[i];
}

static size_t smem_bytes_bwd(const DdObserveBwd& p, int G, int* nsc_out, int* xk_out) {
  const int SC = p.S * p.C;
  const int per = (SC + G - 1) / G;
  const int nsc = (per + NSW - 1) / NSW;
  int xk = max(SC, max(3 * p.D, 2 * p.U));
  xk = (xk + 3) / 4 * 4;
  if (nsc_out) *nsc_out = nsc;
  if (xk_out) *xk_out = xk;
  size_t f = (size_t)SC * NU + (size_t)p.U * NJ + (size_t)3 * p.D * (NJ + NU) +
             (size_t)nsc * p.U * NSW + (size_t)MT * xk;
  return f * sizeof(float);
}

static size_t smem_bytes(const DdObserveFwd& p, int G) {
  const int SC = p.S * p.C, CP = p.C + 4, ngrp = (p.S + G - 1) / G;
  size_t f = (size_t)(SC + p.A) * NU + (size_t)(p.D + p.U) * 3 * NJ + (size_t)p.D * NU +
             (size_t)ngrp * p.U * CP + (size_t)MT * (p.D + p.U) + (size_t)p.S * (1 + MT);
  return f * sizeof(float);
}

}  // namespace scan
}  // namespace dd

using namespace dd;

extern "C" int dd_observe_fwd_grid(const DdObserveFwd* p) {
  // number of CTAs the fused scan needs, or 0 if this shape is not eligible
  if (p->B > scan::MT || p->C > 64 || p->B < 1) return 0;
  const int G = max(max((p->D + scan::NJ - 1) / scan::NJ, (p->U + scan::NU - 1) / scan::NU), 1);
  if (G > 148) return 0;
  if (scan::smem_bytes(*p, G) > 220 * 1024) return 0;
  return G;
}

extern "C" int dd_observe_fwd(const DdObserveFwd* p, dd_stream_t stream) {
  const int G = dd_observe_fwd_grid(p);
  if (G == 0) return fail("dd_observe_fwd", "shape not eligible for the fused scan");
  const size_t smem = scan::smem_bytes(*p, G);
  static bool attr = false;
  if (!attr) {
    cudaFuncSetAttribute(scan::observe_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                         220 * 1024);
    attr = true;
  }
  cudaMemsetAsync(p->barrier, 0, sizeof(unsigned int), S(stream));
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(G);
  cfg.blockDim = dim3(scan::THREADS);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = S(stream);
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeCooperative;       // all CTAs co-resident (grid barriers)
  at[0].val.cooperative = 1;
  cfg.attrs = at;
  cfg.numAttrs = 1;
  DdObserveFwd params = *p;
  cudaError_t e = cudaLaunchKernelEx(&cfg, scan::observe_fwd_kernel, params);
  if (e != cudaSuccess) return fail("dd_observe_fwd", cudaGetErrorString(e));
  return launched("dd_observe_fwd");
}

extern "C" int dd_observe_bwd_grid(const DdObserveBwd* p) {
  if (p->B > scan::MT || p->C > 64 || p->B < 1) return 0;
  const int G = max(max((p->D + scan::NJ - 1) / scan::NJ, (p->U + scan::NU - 1) / scan::NU), 1);
  if (G > 148) return 0;
  if (scan::smem_bytes_bwd(*p, G, nullptr, nullptr) > 220 * 1024) return 0;
  return G;
}

extern "C" int dd_observe_bwd(const DdObserveBwd* p, dd_stream_t stream) {
  const int G = dd_observe_bwd_grid(p);
  if (G == 0) return fail("dd_observe_bwd", "shape not eligible for the fused scan");
  DdObserveBwd params = *p;
  const size_t smem = scan::smem_bytes_bwd(*p, G, &params.nsc, &params.xk);
  static bool attr = false;
  if (!attr) {
    cudaFuncSetAttribute(scan::observe_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                         220 * 1024);
    attr = true;
  }
  cudaMemsetAsync(p->barrier, 0, sizeof(unsigned int), S(stream));
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(G);
  cfg.blockDim = dim3(scan::THREADS);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = S(stream);
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeCooperative;
  at[0].val.cooperative = 1;
  cfg.attrs = at;
  cfg.numAttrs = 1;
  cudaError_t e = cudaLaunchKernelEx(&cfg, scan::observe_bwd_kernel, params);
  if (e != cudaSuccess) return fail("dd_observe_bwd", cudaGetErrorString(e));
  return launched("dd_observe_bwd");
}
