// Image path of the encoder / decoder heads: stride-s VALID convolutions and
// their transposes as (im2col | col2im) around the dense tensor-core GEMM, plus
// the uint8 image staging and the sigmoid-MSE image loss.  NHWC everywhere.
// Reference: embodied/agents/dreamerv2plus/nets.py:298-305 (ImageEncoderSimple),
// :316-327 (ImageDecoderSimple), :495-554 (Conv2D, conv2d_transpose output rule),
// :283-288 (MSEDist over H,W,C); agent.py:123-131 (uint8 / 255).
//
// One column layout serves all four uses: col[(n,ho,wo), (kh,kw,c)] with
// ho = (h - kh) / s.  im2col gathers it from an image (conv forward, conv-transpose
// backward); col2im sums it back into an image (conv backward, conv-transpose
// forward).  HBM-bound streaming kernels; col2im is a gather (no atomics).
#include "dd_common.cuh"

namespace dd {

template <int VEC>
__global__ void im2col_kernel(const float* __restrict__ X, int N, int H, int W, int C, int k,
                              int s, int Ho, int Wo, float* __restrict__ col) {
  pdl_prologue();
  const int Cv = C / VEC;
  const int64_t rowlen = (int64_t)k * k * Cv;
  const int64_t total = (int64_t)N * Ho * Wo * rowlen;
  for (int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; e < total;
       e += (int64_t)gridDim.x * blockDim.x) {
    const int64_t row = e / rowlen;
    const int j = (int)(e - row * rowlen);
    const int cv = j % Cv;
    const int kw = (j / Cv) % k;
    const int kh = j / (Cv * k);
    const int wo = (int)(row % Wo);
    const int ho = (int)((row / Wo) % Ho);
    const int64_t n = row / ((int64_t)Wo * Ho);
    const int h = ho * s + kh, w = wo * s + kw;
    const float* src = X + ((n * H + h) * W + w) * (int64_t)C + (int64_t)cv * VEC;
    float* dst = col + e * VEC;
    if (VEC == 4) {
      *reinterpret_cast<float4*>(dst) = *reinterpret_cast<const float4*>(src);
    } else {
      *dst = *src;
    }
  }
}

template <int VEC>
__global__ void col2im_kernel(const float* __restrict__ col, int N, int H, int W, int C, int k,
                              int s, int Ho, int Wo, const float* __restrict__ bias,
                              float* __restrict__ X) {
  pdl_prologue();
  const int Cv = C / VEC;
  const int64_t total = (int64_t)N * H * W * Cv;
  const int64_t rowlen = (int64_t)k * k * C;
  for (int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; e < total;
       e += (int64_t)gridDim.x * blockDim.x) {
    const int cv = (int)(e % Cv);
    const int w = (int)((e / Cv) % W);
    const int h = (int)((e / ((int64_t)Cv * W)) % H);
    const int64_t n = e / ((int64_t)Cv * W * H);
    float acc[VEC];
#pragma unroll
    for (int v = 0; v < VEC; ++v) acc[v] = bias != nullptr ? bias[cv * VEC + v] : 0.f;
    for (int kh = h % s; kh < k; kh += s) {
      const int ho = (h - kh) / s;
      if (h < kh || ho >= Ho) continue;
      for (int kw = w % s; kw < k; kw += s) {
        const int wo = (w - kw) / s;
        if (w < kw || wo >= Wo) continue;
        const float* src = col + ((n * Ho + ho) * Wo + wo) * rowlen +
                           ((int64_t)kh * k + kw) * C + (int64_t)cv * VEC;
        if (VEC == 4) {
          const float4 t = *reinterpret_cast<const float4*>(src);
          acc[0] += t.x; acc[1] += t.y; acc[2] += t.z; acc[3] += t.w;
        } else {
          acc[0] += *src;
        }
      }
    }
    float* dst = X + e * VEC;
#pragma unroll
    for (int v = 0; v < VEC; ++v) dst[v] = acc[v];
  }
}

// src u8 [B, L, P, Ck] (replay layout) -> dst f32 [L*B, P, Ctot] at channel c0, scaled.
__global__ void image_stage_kernel(const uint8_t* __restrict__ src, int B, int L, int P, int Ck,
                                   float* __restrict__ dst, int Ctot, int c0, float scale) {
  pdl_prologue();
  const int64_t per = (int64_t)P * Ck;
  const int64_t total = (int64_t)B * L * per;
  for (int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; e < total;
       e += (int64_t)gridDim.x * blockDim.x) {
    const int64_t img = e / per;
    const int64_t r = e - img * per;
    const int c = (int)(r % Ck);
    const int64_t p = r / Ck;
    const int b = (int)(img / L), l = (int)(img % L);
    dst[(((int64_t)l * B + b) * P + p) * Ctot + c0 + c] = scale * (float)src[e];
  }
}

// One block per image row m: loss_row[m] = sum_{p, c in [c0,c1)} (sigmoid(z) - t)^2 and
// dZ = coef * 2 (sigmoid(z) - t) sigmoid'(z) on those channels.
__global__ void image_loss_kernel(const float* __restrict__ Z, const float* __restrict__ T, int M,
                                  int P, int Ctot, int c0, int c1, float coef,
                                  float* __restrict__ loss_row, float* __restrict__ dZ) {
  pdl_prologue();
  __shared__ float red[33];
  const int cw = c1 - c0;
  const int64_t n = (int64_t)P * cw;
  for (int m = blockIdx.x; m < M; m += gridDim.x) {
    const int64_t base = (int64_t)m * P * Ctot;
    float part = 0.f;
    for (int64_t e = threadIdx.x; e < n; e += blockDim.x) {
      const int64_t i = base + (e / cw) * Ctot + c0 + (e % cw);
      const float sg = sigmoidf_(Z[i]);
      const float d = sg - T[i];
      part += d * d;
      dZ[i] = coef * 2.f * d * sg * (1.f - sg);
    }
    const float tot = block_sum(part, red);
    if (threadIdx.x == 0) loss_row[m] = tot;
  }
}

__global__ void sigmoid_kernel(const float* __restrict__ x, float* __restrict__ y, int64_t n) {
  pdl_prologue();
  for (int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; e < n;
       e += (int64_t)gridDim.x * blockDim.x)
    y[e] = sigmoidf_(x[e]);
}

static bool conv_dims(int H, int W, int k, int s, int* Ho, int* Wo) {
  if (k <= 0 || s <= 0 || H < k || W < k) return false;
  *Ho = (H - k) / s + 1;
  *Wo = (W - k) / s + 1;
  return true;
}

}  // namespace dd

using namespace dd;

extern "C" int dd_im2col(const float* X, int N, int H, int W, int C, int k, int stride, float* col,
                         dd_stream_t stream) {
  int Ho, Wo;
  if (!conv_dims(H, W, k, stride, &Ho, &Wo)) return fail("dd_im2col", "bad geometry");
  if (N <= 0) return 0;
  const bool v4 = (C % 4 == 0) && ((reinterpret_cast<uintptr_t>(X) | reinterpret_cast<uintptr_t>(col)) & 15) == 0;
  const int64_t total = (int64_t)N * Ho * Wo * k * k * (v4 ? C / 4 : C);
  const int grid = grid_for(total, 256, 148 * 32);
  if (v4) launch_k(im2col_kernel<4>, grid, 256, 0, S(stream), X, N, H, W, C, k, stride, Ho, Wo, col);
  else launch_k(im2col_kernel<1>, grid, 256, 0, S(stream), X, N, H, W, C, k, stride, Ho, Wo, col);
  return launched("dd_im2col");
}

extern "C" int dd_col2im(const float* col, int N, int H, int W, int C, int k, int stride,
                         const float* bias, float* X, dd_stream_t stream) {
  int Ho, Wo;
  if (!conv_dims(H, W, k, stride, &Ho, &Wo)) return fail("dd_col2im", "bad geometry");
  if (N <= 0) return 0;
  const bool v4 = (C % 4 == 0) && ((reinterpret_cast<uintptr_t>(X) | reinterpret_cast<uintptr_t>(col)) & 15) == 0;
  const int64_t total = (int64_t)N * H * W * (v4 ? C / 4 : C);
  const int grid = grid_for(total, 256, 148 * 32);
  if (v4) launch_k(col2im_kernel<4>, grid, 256, 0, S(stream), col, N, H, W, C, k, stride, Ho, Wo, bias, X);
  else launch_k(col2im_kernel<1>, grid, 256, 0, S(stream), col, N, H, W, C, k, stride, Ho, Wo, bias, X);
  return launched("dd_col2im");
}

extern "C" int dd_image_stage(const uint8_t* src, int B, int L, int P, int Ck, float* dst, int Ctot,
                              int c0, float scale, dd_stream_t stream) {
  if (c0 < 0 || c0 + Ck > Ctot) return fail("dd_image_stage", "channel range outside the image");
  const int64_t total = (int64_t)B * L * P * Ck;
  if (total <= 0) return 0;
  launch_k(image_stage_kernel, grid_for(total, 256, 148 * 32), 256, 0, S(stream), src, B, L, P, Ck, dst,
                                                                            Ctot, c0, scale);
  return launched("dd_image_stage");
}

extern "C" int dd_image_loss(const float* Z, const float* target, int M, int P, int Ctot, int c0,
                             int c1, float coef, float* loss_row, float* dZ, dd_stream_t stream) {
  if (c0 < 0 || c1 > Ctot || c1 <= c0) return fail("dd_image_loss", "bad channel range");
  if (M <= 0) return 0;
  launch_k(image_loss_kernel, M < 148 * 8 ? M : 148 * 8, 256, 0, S(stream), Z, target, M, P, Ctot, c0, c1,
                                                                      coef, loss_row, dZ);
  return launched("dd_image_loss");
}

extern "C" int dd_sigmoid(const float* x, float* y, int64_t n, dd_stream_t stream) {
  if (n <= 0) return 0;
  launch_k(sigmoid_kernel, grid_for(n, 256, 148 * 32), 256, 0, S(stream), x, y, n);
  return launched("dd_sigmoid");
}
