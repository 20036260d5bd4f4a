// Reductions, controllers and the optimizer.  Reference:
// embodied/agents/dreamerv2plus/tfutils.py:180-302 (Optimizer: global-norm clip,
// weight decay on 'kernel' tensors, bias-corrected Adam), :414-527 (AutoAdapt,
// Normalize), agent.py:444-454 (slow target).  Controllers live in device
// memory so the whole train step is a static CUDA graph with no host round trip.
#include "dd_common.cuh"

namespace dd {

thread_local char g_err[256] = "";
std::atomic<int64_t> g_launches{0};
int g_pdl = 0;

// ---- moments of a strided 2-D view, accumulated into 8 doubles ------------
__global__ void stats_init_kernel(double* rows, int nrows) {
  pdl_prologue();
  for (int r = blockIdx.x * blockDim.x + threadIdx.x; r < nrows; r += gridDim.x * blockDim.x) {
    double* o = rows + 8 * r;
    o[0] = 0.0; o[1] = 0.0; o[2] = INFINITY; o[3] = -INFINITY;
    o[4] = 0.0; o[5] = 0.0; o[6] = 0.0; o[7] = 0.0;
  }
}

__global__ void stats_kernel(const float* __restrict__ X, int ldx, int rows, int cols,
                             double* __restrict__ out) {
  pdl_prologue();
  __shared__ double sh[6][32];
  const int64_t n = (int64_t)rows * cols;
  double s = 0.0, q = 0.0, a = 0.0;
  float mn = INFINITY, mx = -INFINITY, am = 0.f;
  for (int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; e < n;
       e += (int64_t)gridDim.x * blockDim.x) {
    const float v = X[(size_t)(e / cols) * ldx + (e % cols)];
    s += v;
    q += (double)v * v;
    a += fabsf(v);
    mn = fminf(mn, v);
    mx = fmaxf(mx, v);
    am = fmaxf(am, fabsf(v));
  }
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
  s = warp_sum_d(s);
  q = warp_sum_d(q);
  a = warp_sum_d(a);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    mn = fminf(mn, __shfl_xor_sync(0xffffffffu, mn, o));
    mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    am = fmaxf(am, __shfl_xor_sync(0xffffffffu, am, o));
  }
  if (lane == 0) {
    sh[0][warp] = s; sh[1][warp] = q; sh[2][warp] = mn;
    sh[3][warp] = mx; sh[4][warp] = a; sh[5][warp] = am;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    double ts = 0, tq = 0, ta = 0, tmn = INFINITY, tmx = -INFINITY, tam = 0;
    for (int w = 0; w < nw; ++w) {
      ts += sh[0][w]; tq += sh[1][w]; ta += sh[4][w];
      tmn = fmin(tmn, sh[2][w]); tmx = fmax(tmx, sh[3][w]); tam = fmax(tam, sh[5][w]);
    }
    atomicAdd(&out[0], ts);
    atomicAdd(&out[1], tq);
    atomic_min_d(&out[2], tmn);
    atomic_max_d(&out[3], tmx);
    atomicAdd(&out[4], ta);
    atomic_max_d(&out[5], tam);
    if (blockIdx.x == 0) atomicAdd(&out[6], (double)n);
  }
}

// ---- AutoAdapt 'mult' (tfutils.py:460-474) -----------------------------------
__global__ void autoadapt_kernel(float* scale, const float* sums, int k, float denom,
                                 float target, float lo, float hi, float vel, float thres,
                                 int inverse) {
  pdl_prologue();
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= k) return;
  const float avg = sums[i] / denom;
  bool below = avg < target / (1.f + thres);
  bool above = avg > target * (1.f + thres);
  if (inverse) { const bool t = below; below = above; above = t; }
  float s = scale[i];
  if (above) s = s * (1.f + vel);
  else if (below) s = s / (1.f + vel);
  scale[i] = fminf(fmaxf(s, lo), hi);
}

// ---- Normalize (tfutils.py:502-517), float64 like the reference --------------
__global__ void normalize_update_kernel(double* state, const double* st, double decay,
                                        double maxv, float* out) {
  pdl_prologue();
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  const double n = st[6];
  state[2] += 1.0;
  state[0] = decay * state[0] + (1.0 - decay) * st[0] / n;
  state[1] = decay * state[1] + (1.0 - decay) * st[1] / n;
  const double corr = 1.0 - pow(decay, state[2]);
  const double mean = state[0] / corr;
  const double var = state[1] / corr - mean * mean;
  out[0] = (float)mean;
  out[1] = (float)(1.0 / sqrt(fmax(var, 1.0 / (maxv * maxv))));
}

// ---- optimizer ------------------------------------------------------------------
__global__ void sumsq_kernel(const float* __restrict__ g, int64_t n, double* __restrict__ out) {
  pdl_prologue();
  __shared__ double sh[32];
  double s = 0.0;
  for (int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; e < n;
       e += (int64_t)gridDim.x * blockDim.x) {
    const double v = g[e];
    s += v * v;
  }
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
  s = warp_sum_d(s);
  if (lane == 0) sh[warp] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0.0;
    for (int w = 0; w < nw; ++w) t += sh[w];
    atomicAdd(out, t);
  }
}

__global__ void counter_inc_kernel(int64_t* counters, int ci) {
  pdl_prologue();
  if (threadIdx.x == 0 && blockIdx.x == 0) counters[ci] += 1;
}

__global__ void adam_kernel(float* __restrict__ p, const float* __restrict__ g,
                            float* __restrict__ m, float* __restrict__ v, int64_t n,
                            int64_t n_decay, float lr, double b1d, double b2d, float eps, float wd,
                            float clip, float gscale, const double* __restrict__ sumsq,
                            const int64_t* __restrict__ counters, int ci) {
  pdl_prologue();
  const double t = (double)counters[ci];
  const float b1 = (float)b1d, b2 = (float)b2d;
  const float norm = (float)(sqrt(sumsq[0]) * (double)gscale);
  float s = gscale;
  if (clip > 0.f && norm > 0.f) s *= fminf(1.f, clip / norm);
  const float c1 = (float)(1.0 / (1.0 - pow(b1d, t)));
  const float c2 = (float)(1.0 / (1.0 - pow(b2d, t)));
  const float decay = 1.f - wd * lr;
  for (int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; e < n;
       e += (int64_t)gridDim.x * blockDim.x) {
    const float gg = g[e] * s;
    float pe = p[e];
    if (e < n_decay) pe *= decay;
    const float me = b1 * m[e] + (1.f - b1) * gg;
    const float ve = b2 * v[e] + (1.f - b2) * gg * gg;
    m[e] = me;
    v[e] = ve;
    p[e] = pe - lr * (me * c1) / (sqrtf(ve * c2) + eps);
  }
}

__global__ void slow_copy_kernel(const float* __restrict__ src, float* __restrict__ dst,
                                 int64_t n, const int64_t* __restrict__ counters, int ci,
                                 int every) {
  pdl_prologue();
  const int64_t u = counters[ci];
  if (!(u == -1 || u >= every)) return;
  for (int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; e < n;
       e += (int64_t)gridDim.x * blockDim.x)
    dst[e] = src[e];
}

__global__ void slow_counter_kernel(int64_t* counters, int ci, int every) {
  pdl_prologue();
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  int64_t u = counters[ci];
  if (u == -1 || u >= every) u = 0;
  counters[ci] = u + 1;
}

}  // namespace dd

using namespace dd;

extern "C" int dd_version(void) { return 100; }
extern "C" const char* dd_last_error(void) { return g_err; }
extern "C" int64_t dd_launch_count(void) { return g_launches.load(); }
extern "C" int dd_set_pdl(int on) {
  const int prev = g_pdl;
  g_pdl = on ? 1 : 0;
  return prev;
}

extern "C" int dd_stats_init(double* rows, int nrows, dd_stream_t stream) {
  if (nrows <= 0) return 0;
  launch_k(stats_init_kernel, grid_for(nrows, 64), 64, 0, S(stream), rows, nrows);
  return launched("dd_stats_init");
}

extern "C" int dd_stats(const float* X, int ldx, int rows, int cols, double* out,
                        dd_stream_t stream) {
  const int64_t n = (int64_t)rows * cols;
  if (n <= 0) return 0;
  launch_k(stats_kernel, grid_for(n, 256 * 8, 64), 256, 0, S(stream), X, ldx, rows, cols, out);
  return launched("dd_stats");
}

extern "C" int dd_autoadapt(float* scale, const float* sums, int k, float denom, float target,
                            float lo, float hi, float vel, float thres, int inverse,
                            dd_stream_t stream) {
  if (k <= 0) return 0;
  launch_k(autoadapt_kernel, grid_for(k, 64), 64, 0, S(stream), scale, sums, k, denom, target, lo, hi,
                                                          vel, thres, inverse);
  return launched("dd_autoadapt");
}

extern "C" int dd_normalize_update(double* state, const double* st, double decay, double maxv,
                                   float* out, dd_stream_t stream) {
  launch_k(normalize_update_kernel, 1, 32, 0, S(stream), state, st, decay, maxv, out);
  return launched("dd_normalize_update");
}

extern "C" int dd_sumsq(const float* g, int64_t n, double* out, dd_stream_t stream) {
  if (dd_fill_f64(out, 1, 0.0, stream) != 0) return -1;
  if (n <= 0) return 0;
  launch_k(sumsq_kernel, grid_for(n, 256 * 8, 296), 256, 0, S(stream), g, n, out);
  return launched("dd_sumsq");
}

extern "C" int dd_counter_inc(int64_t* counters, int ci, dd_stream_t stream) {
  launch_k(counter_inc_kernel, 1, 32, 0, S(stream), counters, ci);
  return launched("dd_counter_inc");
}

extern "C" int dd_adam(float* p, const float* g, float* m, float* v, int64_t n, int64_t n_decay,
                       float lr, double b1, double b2, float eps, float wd, float clip,
                       float gscale, const double* sumsq, int64_t* counters, int ci,
                       dd_stream_t stream) {
  if (n <= 0) return 0;
  launch_k(counter_inc_kernel, 1, 32, 0, S(stream), counters, ci);
  launch_k(adam_kernel, grid_for(n, 256 * 4), 256, 0, S(stream), p, g, m, v, n, n_decay, lr, b1, b2, eps,
                                                          wd, clip, gscale, sumsq, counters, ci);
  return launched("dd_adam", 2);
}

extern "C" int dd_slow_sync(const float* src, float* dst, int64_t n, int64_t* counters, int ci,
                            int every, dd_stream_t stream) {
  if (n <= 0) return 0;
  launch_k(slow_copy_kernel, grid_for(n, 256 * 4), 256, 0, S(stream), src, dst, n, counters, ci, every);
  launch_k(slow_counter_kernel, 1, 32, 0, S(stream), counters, ci, every);
  return launched("dd_slow_sync", 2);
}
