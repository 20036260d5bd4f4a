// Elementwise / small-reduction kernels: 2-D helpers, head losses, the actor
// distribution, the lambda-return reverse scan and its backward.  Reference:
// embodied/agents/dreamerv2plus/agent.py:234-261 (imagine: cont, weight),
// :326-381 (actor loss), :398-442 (critic target); tfutils.py:305-356, :395-411;
// nets.py:461-471.  HBM-bound fp32 streaming kernels.
#include "dd_common.cuh"

namespace dd {

#define DD_GRID_STRIDE(e, n)                                                  \
  for (int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; e < (n); \
       e += (int64_t)gridDim.x * blockDim.x)

__global__ void add_kernel(float* dst, int ldd, const float* src, int lds, int rows, int cols) {
  pdl_prologue();
  const int64_t n = (int64_t)rows * cols;
  DD_GRID_STRIDE(e, n) {
    const int r = e / cols, c = e % cols;
    dst[(size_t)r * ldd + c] += src[(size_t)r * lds + c];
  }
}

__global__ void copy_kernel(float* dst, int ldd, const float* src, int lds, int rows, int cols) {
  pdl_prologue();
  const int64_t n = (int64_t)rows * cols;
  DD_GRID_STRIDE(e, n) {
    const int r = e / cols, c = e % cols;
    dst[(size_t)r * ldd + c] = src[(size_t)r * lds + c];
  }
}

__global__ void fill_kernel(float* dst, int ldd, int rows, int cols, float v) {
  pdl_prologue();
  const int64_t n = (int64_t)rows * cols;
  DD_GRID_STRIDE(e, n) {
    const int r = e / cols, c = e % cols;
    dst[(size_t)r * ldd + c] = v;
  }
}

__global__ void fill_f64_kernel(double* dst, int64_t n, double v) {
  pdl_prologue();
  DD_GRID_STRIDE(e, n) dst[e] = v;
}

__global__ void axpy_kernel(float* dst, int ldd, const float* src, int lds, int rows, int cols,
                            const float* scale) {
  pdl_prologue();
  const int64_t n = (int64_t)rows * cols;
  const float a = scale[0];
  DD_GRID_STRIDE(e, n) {
    const int r = e / cols, c = e % cols;
    dst[(size_t)r * ldd + c] += a * src[(size_t)r * lds + c];
  }
}

__global__ void sub_kernel(float* dst, const float* a, const float* b, int64_t n) {
  pdl_prologue();
  DD_GRID_STRIDE(e, n) dst[e] = a[e] - b[e];
}

__global__ void symexp_kernel(const float* src, float* dst, int64_t n) {
  pdl_prologue();
  DD_GRID_STRIDE(e, n) dst[e] = symexpf_(src[e]);
}

__global__ void transpose01_kernel(const float* src, float* dst, int A, int B, int F) {
  pdl_prologue();
  const int64_t n = (int64_t)A * B * F;
  DD_GRID_STRIDE(e, n) {
    const int f = e % F;
    const int64_t ab = e / F;
    const int b = ab % B, a = ab / B;
    dst[((size_t)b * A + a) * F + f] = src[e];
  }
}

// Column sums: block = 32 columns x 8 row-lanes, rows strided over gridDim.y.
__global__ void colsum_kernel(const float* __restrict__ X, int ldx, int M, int N,
                              float* __restrict__ out) {
  pdl_prologue();
  __shared__ float part[8][33];
  const int c = blockIdx.x * 32 + threadIdx.x;
  float acc = 0.f;
  if (c < N) {
    for (int m = blockIdx.y * 8 + threadIdx.y; m < M; m += gridDim.y * 8)
      acc += X[(size_t)m * ldx + c];
  }
  part[threadIdx.y][threadIdx.x] = acc;
  __syncthreads();
  if (threadIdx.y == 0 && c < N) {
    float t = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) t += part[i][threadIdx.x];
    atomicAdd(&out[c], t);
  }
}

__global__ void tanh_fwd_kernel(const float* x, float* y, int64_t n) {
  pdl_prologue();
  DD_GRID_STRIDE(e, n) y[e] = tanhf(x[e]);
}

__global__ void tanh_bwd_kernel(const float* dy, const float* y, float* dx, int64_t n, int acc) {
  pdl_prologue();
  DD_GRID_STRIDE(e, n) {
    const float g = dy[e] * (1.f - y[e] * y[e]);
    dx[e] = acc ? dx[e] + g : g;
  }
}

// ---------------------------------------------------------------------------
// Head losses: one thread per row (dim is small: 1 .. a few dozen).
// ---------------------------------------------------------------------------
__global__ void head_loss_kernel(int kind, const float* __restrict__ pred, int ldp,
                                 const float* __restrict__ target, int ldt,
                                 const float* __restrict__ weight, float coef,
                                 float* __restrict__ loss_row, float* __restrict__ dpred,
                                 int lddp, int M, int dim) {
  pdl_prologue();
  DD_GRID_STRIDE(m, M) {
    const float w = weight != nullptr ? weight[m] : 1.f;
    float loss = 0.f;
    for (int d = 0; d < dim; ++d) {
      const float p = pred[(size_t)m * ldp + d];
      const float t = target[(size_t)m * ldt + d];
      float l, g;
      if (kind == 0) {
        const float diff = p - t;
        l = diff * diff;
        g = 2.f * diff;
      } else if (kind == 1) {
        const float diff = p - symlogf_(t);
        l = diff * diff;
        g = 2.f * diff;
      } else {
        // softplus(p) - t*p, stable
        l = fmaxf(p, 0.f) + log1pf(expf(-fabsf(p))) - t * p;
        g = sigmoidf_(p) - t;
      }
      loss += l;
      dpred[(size_t)m * lddp + d] = coef * w * g;
    }
    loss_row[m] = w * loss;
  }
}

// balance_stats sums (single CTA; M is the replay batch, a few thousand rows)
__global__ void balance_kernel(int kind, const float* __restrict__ pred,
                               const float* __restrict__ loss_row,
                               const float* __restrict__ target, float thres, int M,
                               double* __restrict__ out) {
  pdl_prologue();
  __shared__ double sh[8][32];
  double v[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  for (int m = threadIdx.x; m < M; m += blockDim.x) {
    const float mean = kind == 1 ? symexpf_(pred[m]) : sigmoidf_(pred[m]);
    const float t = target[m];
    const double pos = t > thres ? 1.0 : 0.0, neg = 1.0 - pos;
    const double pr = mean > thres ? 1.0 : 0.0;
    const double l = loss_row[m];
    v[0] += pos; v[1] += neg; v[2] += l * pos; v[3] += l * neg;
    v[4] += pr * pos; v[5] += (1.0 - pr) * neg; v[6] += t; v[7] += mean;
  }
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    v[i] = warp_sum_d(v[i]);
    if (lane == 0) sh[i][warp] = v[i];
  }
  __syncthreads();
  if (threadIdx.x < 8) {
    double t = 0.0;
    for (int w = 0; w < (blockDim.x >> 5); ++w) t += sh[threadIdx.x][w];
    out[threadIdx.x] = t;
  }
}

// ---------------------------------------------------------------------------
// Actor distribution (continuous 'normal' head, nets.py:461-468)
// ---------------------------------------------------------------------------
__global__ void actor_sample_kernel(const float* mraw, const float* sraw, const float* eps,
                                    float lo, float hi, float* action, float* mean, float* std,
                                    int64_t n) {
  pdl_prologue();
  DD_GRID_STRIDE(e, n) {
    const float mu = tanhf(mraw[e]);
    const float sd = (hi - lo) * sigmoidf_(sraw[e]) + lo;
    mean[e] = mu;
    std[e] = sd;
    action[e] = mu + sd * eps[e];
  }
}

__global__ void actor_ent_kernel(const float* std, float lo, float hi, float* ent, int64_t n) {
  pdl_prologue();
  const float inv = 1.f / (logf(hi) - logf(lo));
  const float llo = logf(lo);
  DD_GRID_STRIDE(e, n) ent[e] = (logf(std[e]) - llo) * inv;
}

__global__ void actor_loss_kernel(const float* ret, const float* value, const float* adv,
                                  const float* ent, const float* scale, const float* weight,
                                  float* out, int64_t M, int A) {
  pdl_prologue();
  const float mean = adv[0], sc = adv[1];
  DD_GRID_STRIDE(m, M) {
    const float score = (ret[m] - value[m] - mean) * sc;
    float e = 0.f;
    for (int a = 0; a < A; ++a) e += ent[m * A + a] * scale[a];
    out[m] = (-score - e) * weight[m];
  }
}

__global__ void actor_bwd_kernel(const float* dA, const float* eps, const float* mean,
                                 const float* std, const float* weight, const float* scale,
                                 float coef, float lo, float hi, float* dmraw, float* dsraw,
                                 int64_t M, int A) {
  pdl_prologue();
  const float inv = 1.f / (logf(hi) - logf(lo));
  const int64_t n = M * A;
  DD_GRID_STRIDE(e, n) {
    const int64_t m = e / A;
    const int a = e % A;
    const float sd = std[e], mu = mean[e], g = dA[e];
    const float dstd = g * eps[e] - coef * weight[m] * scale[a] * inv / sd;
    dmraw[e] = g * (1.f - mu * mu);
    const float sg = (sd - lo) / (hi - lo);
    dsraw[e] = dstd * (hi - lo) * sg * (1.f - sg);
  }
}

// ---------------------------------------------------------------------------
// lambda-return 'gve' (agent.py:434-440): one thread per imagined start row,
// reverse scan over the horizon; cont/weight of agent.py:256-259 fused in.
// ---------------------------------------------------------------------------
__global__ void lambda_fwd_kernel(const float* __restrict__ rho, const float* __restrict__ vraw,
                                  const float* __restrict__ kraw,
                                  const float* __restrict__ cont0, float gamma, float lam,
                                  float* __restrict__ reward, float* __restrict__ value,
                                  float* __restrict__ cont, float* __restrict__ weight,
                                  float* __restrict__ ret, int H, int N) {
  pdl_prologue();
  DD_GRID_STRIDE(n, N) {
    float w = 1.f / gamma;
    for (int t = 0; t <= H; ++t) {
      const size_t i = (size_t)t * N + n;
      const float c = t == 0 ? cont0[n] : sigmoidf_(kraw[i]);
      cont[i] = c;
      w *= gamma * c;
      weight[i] = w;
      value[i] = symexpf_(vraw[i]);
      if (t > 0) reward[(size_t)(t - 1) * N + n] = symexpf_(rho[i]);
    }
    float last = value[(size_t)H * N + n];
    for (int t = H - 1; t >= 0; --t) {
      const size_t i1 = (size_t)(t + 1) * N + n;
      const float d = gamma * cont[i1];
      last = reward[(size_t)t * N + n] + d * ((1.f - lam) * value[i1] + lam * last);
      ret[(size_t)t * N + n] = last;
    }
  }
}

__global__ void lambda_bwd_kernel(const float* __restrict__ rho, const float* __restrict__ vraw,
                                  const float* __restrict__ value,
                                  const float* __restrict__ cont,
                                  const float* __restrict__ weight,
                                  const float* __restrict__ ret, const float* __restrict__ adv,
                                  float coef, float gamma, float lam, float* __restrict__ drho,
                                  float* __restrict__ dvraw, float* __restrict__ dkraw, int H,
                                  int N) {
  pdl_prologue();
  const float sc = adv[1];
  DD_GRID_STRIDE(n, N) {
    float g = 0.f;
    // t = 0: only the baseline term touches V_0
    {
      const float dret0 = -coef * sc * weight[n];
      dvraw[n] = -dret0 * expf(fabsf(vraw[n]));
      drho[n] = 0.f;
      dkraw[n] = 0.f;
    }
    for (int t = 1; t <= H; ++t) {
      const size_t i = (size_t)t * N + n, ip = (size_t)(t - 1) * N + n;
      const float dret_prev = -coef * sc * weight[ip];
      g = dret_prev + (t > 1 ? g * gamma * cont[ip] * lam : 0.f);
      const float c = cont[i];
      const float d = gamma * c;
      const float nxt = t < H ? ret[i] : value[i];
      const float dr = g;
      const float ddisc = g * ((1.f - lam) * value[i] + lam * nxt);
      float dV = g * d * ((1.f - lam) + (t == H ? lam : 0.f));
      if (t < H) dV += coef * sc * weight[i];          // baseline: -dret_t
      drho[i] = dr * expf(fabsf(rho[i]));
      dvraw[i] = dV * expf(fabsf(vraw[i]));
      dkraw[i] = ddisc * gamma * c * (1.f - c);
    }
  }
}

// ---------------------------------------------------------------------------
// Discrete actor ('onehot' head with 'reinforce' gradient, nets.py:480-491,
// agent.py:351-381).  One warp per row, A <= 64 classes.
// ---------------------------------------------------------------------------
__device__ __forceinline__ void row_log_softmax(const float* lg, int A, int lane, float a[2]) {
  const float l0 = lane < A ? lg[lane] : -INFINITY;
  const float l1 = lane + 32 < A ? lg[lane + 32] : -INFINITY;
  const float mx = warp_max(fmaxf(l0, l1));
  const float se = warp_sum((lane < A ? expf(l0 - mx) : 0.f) + (lane + 32 < A ? expf(l1 - mx) : 0.f));
  const float lse = mx + logf(se);
  a[0] = l0 - lse;
  a[1] = l1 - lse;
}

__global__ void actor_disc_ent_kernel(const float* __restrict__ logit, float* __restrict__ ent,
                                      int64_t M, int A) {
  pdl_prologue();
  const int lane = threadIdx.x & 31, wpb = blockDim.x >> 5;
  const float inv = 1.f / logf((float)A);
  for (int64_t m = blockIdx.x * (int64_t)wpb + (threadIdx.x >> 5); m < M;
       m += (int64_t)gridDim.x * wpb) {
    float a[2];
    row_log_softmax(logit + m * A, A, lane, a);
    float h = 0.f;
    if (lane < A) h -= expf(a[0]) * a[0];
    if (lane + 32 < A) h -= expf(a[1]) * a[1];
    h = warp_sum(h);
    if (lane == 0) ent[m] = h * inv;
  }
}

__global__ void actor_disc_loss_kernel(const float* __restrict__ logit,
                                       const float* __restrict__ action,
                                       const float* __restrict__ ret,
                                       const float* __restrict__ value,
                                       const float* __restrict__ adv,
                                       const float* __restrict__ ent,
                                       const float* __restrict__ scale,
                                       const float* __restrict__ weight, float coef,
                                       float* __restrict__ loss_row, float* __restrict__ dlogit,
                                       int64_t M, int A) {
  pdl_prologue();
  const int lane = threadIdx.x & 31, wpb = blockDim.x >> 5;
  const float lnA = logf((float)A);
  const float mean = adv[0], sc = adv[1], es = scale[0];
  for (int64_t m = blockIdx.x * (int64_t)wpb + (threadIdx.x >> 5); m < M;
       m += (int64_t)gridDim.x * wpb) {
    float a[2];
    row_log_softmax(logit + m * A, A, lane, a);
    const float* act = action + m * A;
    const float x0 = lane < A ? act[lane] : 0.f;
    const float x1 = lane + 32 < A ? act[lane + 32] : 0.f;
    const float q0 = lane < A ? expf(a[0]) : 0.f;
    const float q1 = lane + 32 < A ? expf(a[1]) : 0.f;
    float logp = (lane < A ? x0 * a[0] : 0.f) + (lane + 32 < A ? x1 * a[1] : 0.f);
    float h = -((lane < A ? q0 * a[0] : 0.f) + (lane + 32 < A ? q1 * a[1] : 0.f));
    float xs = x0 + x1;
    logp = warp_sum(logp);
    h = warp_sum(h);
    xs = warp_sum(xs);
    const float score = (ret[m] - value[m] - mean) * sc;
    const float w = weight[m];
    if (lane == 0) loss_row[m] = (-logp * score - es * ent[m]) * w;
    const float k = coef * w;
    float* dl = dlogit + m * A;
    if (lane < A) dl[lane] = k * (score * (q0 * xs - x0) + es / lnA * q0 * (a[0] + h));
    if (lane + 32 < A) dl[lane + 32] = k * (score * (q1 * xs - x1) + es / lnA * q1 * (a[1] + h));
  }
}

}  // namespace dd

using namespace dd;

#define DD_EW(kernel, n, ...)                                                     \
  launch_k(kernel, grid_for((n), 256), 256, 0, S(stream), __VA_ARGS__)

extern "C" int dd_add(float* dst, int ldd, const float* src, int lds, int rows, int cols,
                      dd_stream_t stream) {
  if ((int64_t)rows * cols <= 0) return 0;
  DD_EW(add_kernel, (int64_t)rows * cols, dst, ldd, src, lds, rows, cols);
  return launched("dd_add");
}

extern "C" int dd_copy(float* dst, int ldd, const float* src, int lds, int rows, int cols,
                       dd_stream_t stream) {
  if ((int64_t)rows * cols <= 0) return 0;
  DD_EW(copy_kernel, (int64_t)rows * cols, dst, ldd, src, lds, rows, cols);
  return launched("dd_copy");
}

extern "C" int dd_fill(float* dst, int ldd, int rows, int cols, float value, dd_stream_t stream) {
  if ((int64_t)rows * cols <= 0) return 0;
  DD_EW(fill_kernel, (int64_t)rows * cols, dst, ldd, rows, cols, value);
  return launched("dd_fill");
}

extern "C" int dd_fill_f64(double* dst, int64_t n, double value, dd_stream_t stream) {
  if (n <= 0) return 0;
  DD_EW(fill_f64_kernel, n, dst, n, value);
  return launched("dd_fill_f64");
}

extern "C" int dd_axpy(float* dst, int ldd, const float* src, int lds, int rows, int cols,
                       const float* scale, dd_stream_t stream) {
  if ((int64_t)rows * cols <= 0) return 0;
  DD_EW(axpy_kernel, (int64_t)rows * cols, dst, ldd, src, lds, rows, cols, scale);
  return launched("dd_axpy");
}

extern "C" int dd_sub(float* dst, const float* a, const float* b, int64_t n, dd_stream_t stream) {
  if (n <= 0) return 0;
  DD_EW(sub_kernel, n, dst, a, b, n);
  return launched("dd_sub");
}

extern "C" int dd_symexp(const float* src, float* dst, int64_t n, dd_stream_t stream) {
  if (n <= 0) return 0;
  DD_EW(symexp_kernel, n, src, dst, n);
  return launched("dd_symexp");
}

extern "C" int dd_transpose01(const float* src, float* dst, int A, int B, int F,
                              dd_stream_t stream) {
  const int64_t n = (int64_t)A * B * F;
  if (n <= 0) return 0;
  DD_EW(transpose01_kernel, n, src, dst, A, B, F);
  return launched("dd_transpose01");
}

extern "C" int dd_colsum(const float* X, int ldx, int M, int N, float* out, int acc,
                         dd_stream_t stream) {
  if (N <= 0) return 0;
  int nl = 0;
  if (!acc) {
    launch_k(fill_kernel, grid_for(N, 256), 256, 0, S(stream), out, N, 1, N, 0.f);
    ++nl;
  }
  if (M > 0) {
    int gy = (M + 63) / 64;
    if (gy > 296) gy = 296;
    dim3 grid((N + 31) / 32, gy), block(32, 8);
    launch_k(colsum_kernel, grid, block, 0, S(stream), X, ldx, M, N, out);
    ++nl;
  }
  return launched("dd_colsum", nl);
}

extern "C" int dd_tanh_fwd(const float* x, float* y, int64_t n, dd_stream_t stream) {
  if (n <= 0) return 0;
  DD_EW(tanh_fwd_kernel, n, x, y, n);
  return launched("dd_tanh_fwd");
}

extern "C" int dd_tanh_bwd(const float* dy, const float* y, float* dx, int64_t n, int acc,
                           dd_stream_t stream) {
  if (n <= 0) return 0;
  DD_EW(tanh_bwd_kernel, n, dy, y, dx, n, acc);
  return launched("dd_tanh_bwd");
}

extern "C" int dd_head_loss(int kind, const float* pred, int ldp, const float* target, int ldt,
                            const float* weight, float coef, float* loss_row, float* dpred,
                            int lddp, int M, int dim, dd_stream_t stream) {
  if (M <= 0) return 0;
  if (kind < 0 || kind > 2) return fail("dd_head_loss", "kind must be 0, 1 or 2");
  launch_k(head_loss_kernel, grid_for(M, 128), 128, 0, S(stream), kind, pred, ldp, target, ldt, weight,
                                                            coef, loss_row, dpred, lddp, M, dim);
  return launched("dd_head_loss");
}

extern "C" int dd_balance(int kind, const float* pred, const float* loss_row, const float* target,
                          float thres, int M, double* out, dd_stream_t stream) {
  launch_k(balance_kernel, 1, 256, 0, S(stream), kind, pred, loss_row, target, thres, M, out);
  return launched("dd_balance");
}

extern "C" int dd_actor_sample(const float* mraw, const float* sraw, const float* eps, float lo,
                               float hi, float* action, float* mean, float* std, int64_t n,
                               dd_stream_t stream) {
  if (n <= 0) return 0;
  DD_EW(actor_sample_kernel, n, mraw, sraw, eps, lo, hi, action, mean, std, n);
  return launched("dd_actor_sample");
}

extern "C" int dd_actor_ent(const float* std, float lo, float hi, float* ent, int64_t n,
                            dd_stream_t stream) {
  if (n <= 0) return 0;
  DD_EW(actor_ent_kernel, n, std, lo, hi, ent, n);
  return launched("dd_actor_ent");
}

extern "C" int dd_actor_loss(const float* ret, const float* value, const float* adv,
                             const float* ent, const float* scale, const float* weight, float* out,
                             int64_t M, int A, dd_stream_t stream) {
  if (M <= 0) return 0;
  DD_EW(actor_loss_kernel, M, ret, value, adv, ent, scale, weight, out, M, A);
  return launched("dd_actor_loss");
}

extern "C" int dd_actor_bwd(const float* dA, const float* eps, const float* mean, const float* std,
                            const float* weight, const float* scale, float coef, float lo, float hi,
                            float* dmraw, float* dsraw, int64_t M, int A, dd_stream_t stream) {
  if (M <= 0) return 0;
  DD_EW(actor_bwd_kernel, M * A, dA, eps, mean, std, weight, scale, coef, lo, hi, dmraw, dsraw, M,
        A);
  return launched("dd_actor_bwd");
}

extern "C" int dd_actor_disc_ent(const float* logit, float* ent, int64_t M, int A,
                                 dd_stream_t stream) {
  if (M <= 0) return 0;
  if (A > 64 || A < 2) return fail("dd_actor_disc_ent", "needs 2 <= classes <= 64");
  launch_k(actor_disc_ent_kernel, grid_for(M, 8), 256, 0, S(stream), logit, ent, M, A);
  return launched("dd_actor_disc_ent");
}

extern "C" int dd_actor_disc_loss(const float* logit, const float* action, const float* ret,
                                  const float* value, const float* adv, const float* ent,
                                  const float* scale, const float* weight, float coef,
                                  float* loss_row, float* dlogit, int64_t M, int A,
                                  dd_stream_t stream) {
  if (M <= 0) return 0;
  if (A > 64 || A < 2) return fail("dd_actor_disc_loss", "needs 2 <= classes <= 64");
  launch_k(actor_disc_loss_kernel, grid_for(M, 8), 256, 0, S(stream), 
      logit, action, ret, value, adv, ent, scale, weight, coef, loss_row, dlogit, M, A);
  return launched("dd_actor_disc_loss");
}

extern "C" int dd_lambda_fwd(const float* rho, const float* vraw, const float* kraw,
                             const float* cont0, float gamma, float lam, float* reward,
                             float* value, float* cont, float* weight, float* ret, int H, int N,
                             dd_stream_t stream) {
  if (N <= 0) return 0;
  launch_k(lambda_fwd_kernel, grid_for(N, 64), 64, 0, S(stream), rho, vraw, kraw, cont0, gamma, lam,
                                                           reward, value, cont, weight, ret, H, N);
  return launched("dd_lambda_fwd");
}

extern "C" int dd_lambda_bwd(const float* rho, const float* vraw, const float* value,
                             const float* cont, const float* weight, const float* ret,
                             const float* adv, float coef, float gamma, float lam, float* drho,
                             float* dvraw, float* dkraw, int H, int N, dd_stream_t stream) {
  if (N <= 0) return 0;
  launch_k(lambda_bwd_kernel, grid_for(N, 64), 64, 0, S(stream), rho, vraw, value, cont, weight, ret,
                                                           adv, coef, gamma, lam, drho, dvraw,
                                                           dkraw, H, N);
  return launched("dd_lambda_bwd");
}
