/* dd_kernels.h -- C-ABI of the B200 (sm_100a) kernel library behind the
 * DreamerV2+ learner hot path.
 *
 * The reference (danijar/daydreamer) has no FFI on this path: its learner calls
 * TensorFlow ops from Python (embodied/agents/dreamerv2plus/{agent,nets,tfutils}.py).
 * This ABI is therefore the seam a maintainer would bind with ctypes from the
 * reference-side Python; every entry point names the
 * reference code whose arithmetic it replaces.  A/ = embodied/agents/dreamerv2plus.
 *
 * Conventions
 *   - every pointer is a caller-owned DEVICE pointer (fp32 unless stated);
 *     matrices are row-major with an explicit row stride `ld*` in elements;
 *   - no allocation, no host synchronisation, no global state: each call only
 *     enqueues kernels on `stream` (cudaStream_t passed as void*), so a whole
 *     train step can be captured into one CUDA graph;
 *   - return value 0 = enqueued, negative = rejected argument / launch error,
 *     message via dd_last_error() (thread-local).
 */
#ifndef DD_KERNELS_H_
#define DD_KERNELS_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef void* dd_stream_t;

int dd_version(void);
const char* dd_last_error(void);
/* number of kernel launches enqueued by this library since process start */
int64_t dd_launch_count(void);
/* launch every kernel with programmatic stream serialization (PDL): consecutive
 * launches overlap their launch latency, data dependences are unchanged.
 * Returns the previous setting. */
int dd_set_pdl(int on);

/* ---- dense ------------------------------------------------------------ */
/* C[M,N] = op(A) op(B) (+ bias[N]) + beta*C ; op = transpose when t*=1.
 * Replaces `x @ kernel (+ bias)` of Linear (A/nets.py:566-579) and the
 * matmul gradients TF derives for it.  beta must be 0 or 1. */
int dd_gemm(const float* A, int lda, int ta, const float* B, int ldb, int tb,
            float* C, int ldc, int M, int N, int K, float beta,
            const float* bias, dd_stream_t stream);

/* 1 (default): eligible shapes run on the tcgen05 3xTF32 kernel; 0: always the
 * fp32 SIMT kernel.  Returns the previous setting.  (env DD_GEMM_TC=0 too.) */
int dd_set_gemm_tc(int on);
/* 1 (default): GEMMs with more 128x128 tiles than ~1.5x the SM count use the persistent
 * tensor-core kernel (double-buffered TMEM, overlapped epilogue).  Returns previous. */
int dd_set_gemm_tc_persistent(int on);
/* EXPERIMENTAL (round 2, not yet run on a GPU): cp.async-staged operand loader for the 64-wide
 * tensor-core tile, four k-blocks of L2 latency in flight instead of two (default off) */
int dd_set_gemm_tc_staged(int on);
/* thin GEMMs (N <= 32 or K <= 32 with many rows) on dedicated streaming kernels (default off) */
int dd_set_gemm_thin(int on);
/* widest row the warp-per-row LayerNorm kernels take (<= 512, default 128) */
int dd_set_ln_warp_max(int n);
/* EXPERIMENTAL (round 2, not yet run on a GPU): float4 warp-per-row LayerNorm kernels for
 * N % 128 == 0, N <= 512 (default off) */
int dd_set_ln_vec(int on);
/* profiling aid: bit 0 = skip operand loads, bit 1 = skip MMA issue (garbage results) */
int dd_set_gemm_tc_debug(int bits);

/* LayerNorm(eps 1e-3, population variance) + optional elu (A/nets.py:585-602,
 * :580-581).  stats[M,2] = (mean, rstd).  act: 0 none, 1 elu. */
int dd_ln_fwd(const float* Z, int ldz, const float* scale, const float* bias,
              float* H, int ldh, float* stats, int M, int N, int act,
              dd_stream_t stream);
/* dZ from dH; dscale/dbias (+=) may be NULL to skip parameter gradients. */
int dd_ln_bwd(const float* dH, int lddh, const float* Z, int ldz,
              const float* stats, const float* scale, const float* bias,
              float* dZ, int lddz, float* dscale, float* dbias, int M, int N,
              int act, dd_stream_t stream);

/* GRU gates on the post-LayerNorm 3D vector (A/nets.py:155-160):
 * r=sig(y_r); c=tanh(r*y_c); u=sig(y_u-1); out=u*c+(1-u)*prev. */
int dd_gru_fwd(const float* Y, int ldy, const float* prev, int ldp, float* out,
               int ldo, int M, int D, dd_stream_t stream);
int dd_gru_bwd(const float* dOut, int lddo, const float* Y, int ldy,
               const float* prev, int ldp, float* dY, int lddy, float* dPrev,
               int lddp, int M, int D, int acc_prev, dd_stream_t stream);

/* ---- categorical latent ------------------------------------------------ */
/* unimix logits log((1-mix) softmax(z) + mix/C) (A/nets.py:162-171) and the
 * OneHotDist draw idx = argmax(logit + G) (A/tfutils.py:368-382).
 * mode 0: sample with Gumbel noise G; 1: argmax (dist.mode, A/nets.py:147);
 * 2: logits only.  idx (int32 [M,S]) may be NULL. */
int dd_stats_fwd(const float* Z, int ldz, const float* G, int ldg, float* logit,
                 int ldl, float* stoch, int lds, int32_t* idx, int M, int S,
                 int C, float mix, int mode, dd_stream_t stream);
/* straight-through gradient (dStoch, wrt probs) + direct logit gradient
 * (dLogit), either may be NULL, pulled back to the raw linear output. */
int dd_stats_bwd(const float* dStoch, int ldds, const float* dLogit, int lddl,
                 const float* logit, int ldl, float* dZ, int lddz, int M, int S,
                 int C, float mix, dd_stream_t stream);
/* KL(post||prior) summed over S groups + entropies (A/nets.py:178-183,
 * A/agent.py:198-201). */
int dd_kl_fwd(const float* post, int ldp, const float* prior, int ldq, int M,
              int S, int C, float* kl_row, float* ent_post, float* ent_prior,
              dd_stream_t stream);
/* gradients of coef*scale[0]*(bal*KL(sg(post)||prior)+(1-bal)*KL(post||sg(prior))) */
int dd_kl_bwd(const float* post, int ldp, const float* prior, int ldq, int M,
              int S, int C, float bal, float coef, const float* scale,
              float* dpost, int lddp, float* dprior, int lddq,
              dd_stream_t stream);

/* ---- fused observe scan (forward) ---------------------------------------- */
/* RSSM.observe (A/nets.py:66-76): all L obs_steps (A/nets.py:99-117) in ONE
 * persistent cooperative kernel.  Time-major buffers, row = t*B + b.
 * z_obs must be pre-filled with embed @ obs_out.kernel[D:] (the batched part).
 * Semantics per step are exactly dd_mask_fwd -> dd_gemm/dd_ln_fwd (img_in) ->
 * dd_gemm/dd_ln_fwd/dd_gru_fwd (GRU) -> dd_gemm/dd_ln_fwd (obs_out) ->
 * dd_gemm + dd_stats_fwd mode 0 (obs_stats). */
typedef struct DdObserveFwd {
  int B, L, D, U, S, C, A;
  float mix;
  const float *action, *first, *first0, *prev_deter, *prev_stoch;
  const float *init_deter, *init_stoch, *g_obs;
  const float *Win, *ln_in_s, *ln_in_b, *Wgru, *ln_gru_s, *ln_gru_b;
  const float *Wobs, *ln_obs_s, *ln_obs_b, *Wst, *bst;
  float *m_deter, *m_stoch, *m_act, *z_in, *x_in, *st_in, *z_gru, *y_gru, *st_gru;
  float *feat, *z_obs, *h_obs, *st_obs, *post_logit;
  int32_t* post_idx;
  float* partial;          /* workspace: 148 * 16 * 2 floats */
  unsigned int* barrier;   /* workspace: 1 counter */
  long long* timing;       /* optional: 10 per-phase SM-clock totals of CTA 0, or NULL */
} DdObserveFwd;
/* CTAs the fused scan would use, or 0 when the shape is not eligible
 * (B > 16, weights do not fit the shared-memory slices, ...). */
int dd_observe_fwd_grid(const DdObserveFwd* p);
int dd_observe_fwd(const DdObserveFwd* p, dd_stream_t stream);

/* ---- fused observe scan, backward (BPTT) --------------------------------
 * One persistent cooperative kernel for all L reverse steps; per step exactly
 * dd_stats_bwd -> dd_gemm(tb) -> dd_ln_bwd -> dd_gemm(tb) -> dd_gru_bwd ->
 * dd_ln_bwd (joint 3D) -> 2x dd_gemm(tb) -> dd_ln_bwd -> dd_gemm(tb) -> dd_mask_bwd
 * without the LayerNorm parameter gradients (the caller gets them from dh / dy_gru /
 * dx with three batched dd_ln_bwd calls).  dfeat [L*B, D+SC] holds the head / prior
 * gradients on entry and receives the carried gradients in place. */
typedef struct DdObserveBwd {
  int B, L, D, U, S, C;
  int nsc, xk;             /* filled by the library */
  float mix;
  const float *first, *first0, *post_logit, *dpost_logit;
  const float *z_obs, *st_obs, *ln_obs_s, *ln_obs_b;
  const float *y_gru, *z_gru, *st_gru, *ln_gru_s, *ln_gru_b, *m_deter;
  const float *z_in, *st_in, *ln_in_s, *ln_in_b;
  const float *Wst, *Wobs, *Wgru, *Win;
  float *dfeat, *dzs, *dz_obs, *dz_gru, *dz_in, *dh, *dy_gru, *dx, *d_initdeter;
  float* partial;          /* workspace: 148 * 16 * 2 floats */
  unsigned int* barrier;   /* workspace: 1 counter */
  long long* timing;       /* optional: 12 per-phase SM-clock totals of CTA 0, or NULL */
} DdObserveBwd;
int dd_observe_bwd_grid(const DdObserveBwd* p);
int dd_observe_bwd(const DdObserveBwd* p, dd_stream_t stream);

/* ---- observe-scan masking (A/nets.py:100-107) ------------------------ */
int dd_mask_fwd(const float* pd, int ldpd, const float* ps, int ldps,
                const float* act, int lda, const float* first,
                const float* init_d, const float* init_s, float* od, int ldod,
                float* os, int ldos, float* oa, int ldoa, int M, int D, int SC,
                int A, dd_stream_t stream);
/* dpd/dps (+=) may be NULL (first step of a chunk); dinit[D] += sum_m dd*first */
int dd_mask_bwd(const float* dd, int lddd, const float* ds, int ldds,
                const float* first, float* dpd, int lddpd, float* dps,
                int lddps, float* dinit, int M, int D, int SC,
                dd_stream_t stream);

/* ---- 2-D elementwise helpers ------------------------------------------ */
int dd_add(float* dst, int ldd, const float* src, int lds, int rows, int cols,
           dd_stream_t stream);                       /* dst += src */
int dd_copy(float* dst, int ldd, const float* src, int lds, int rows, int cols,
            dd_stream_t stream);
int dd_fill(float* dst, int ldd, int rows, int cols, float value,
            dd_stream_t stream);
int dd_fill_f64(double* dst, int64_t n, double value, dd_stream_t stream);
int dd_axpy(float* dst, int ldd, const float* src, int lds, int rows, int cols,
            const float* scale, dd_stream_t stream);  /* dst += scale[0]*src */
int dd_sub(float* dst, const float* a, const float* b, int64_t n,
           dd_stream_t stream);
int dd_symexp(const float* src, float* dst, int64_t n, dd_stream_t stream);
int dd_transpose01(const float* src, float* dst, int A, int B, int F,
                   dd_stream_t stream);               /* [A,B,F] -> [B,A,F] */
int dd_colsum(const float* X, int ldx, int M, int N, float* out, int acc,
              dd_stream_t stream);
int dd_tanh_fwd(const float* x, float* y, int64_t n, dd_stream_t stream);
int dd_tanh_bwd(const float* dy, const float* y, float* dx, int64_t n, int acc,
                dd_stream_t stream);

/* ---- image path: convolutions as (im2col | col2im) + dd_gemm ---------------
 * NHWC; col[(n,ho,wo),(kh,kw,c)], ho=(H-k)/stride+1 (VALID).  im2col serves the
 * conv forward (A/nets.py:541-547) and the conv-transpose backward; col2im (+bias)
 * the conv backward and the conv-transpose forward (A/nets.py:523-539). */
int dd_im2col(const float* X, int N, int H, int W, int C, int k, int stride,
              float* col, dd_stream_t stream);
int dd_col2im(const float* col, int N, int H, int W, int C, int k, int stride,
              const float* bias, float* X, dd_stream_t stream);
/* replay uint8 [B,L,P,Ck] -> f32 time-major [L*B,P,Ctot] at channel c0, times
 * scale (1/255; A/agent.py:128-129) -- channel concat of A/nets.py:218. */
int dd_image_stage(const uint8_t* src, int B, int L, int P, int Ck, float* dst,
                   int Ctot, int c0, float scale, dd_stream_t stream);
/* MSEDist(sigmoid(Z), 3, 'sum') on channels [c0,c1) (A/nets.py:283-288,325):
 * loss_row[m] = sum (sigmoid(z)-t)^2 ; dZ = coef * dloss/dz on those channels. */
int dd_image_loss(const float* Z, const float* target, int M, int P, int Ctot,
                  int c0, int c1, float coef, float* loss_row, float* dZ,
                  dd_stream_t stream);
int dd_sigmoid(const float* x, float* y, int64_t n, dd_stream_t stream);

/* ---- losses ------------------------------------------------------------ */
/* kind 0: MSEDist sum (A/tfutils.py:320-329); 1: SymlogDist (:347-356);
 * 2: -Bernoulli(logits).log_prob (A/nets.py:469-471).  weight[M] may be NULL.
 * loss_row[m] = w*loss ; dpred = coef*w*dloss/dpred. */
int dd_head_loss(int kind, const float* pred, int ldp, const float* target,
                 int ldt, const float* weight, float coef, float* loss_row,
                 float* dpred, int lddp, int M, int dim, dd_stream_t stream);
/* balance_stats sums (A/tfutils.py:395-411); out f64[8]. */
int dd_balance(int kind, const float* pred, const float* loss_row,
               const float* target, float thres, int M, double* out,
               dd_stream_t stream);

/* ---- imagination / actor-critic ---------------------------------------- */
/* mean=tanh(m), std=(hi-lo)sigmoid(s)+lo, a=mean+std*eps (A/nets.py:461-468). */
int dd_actor_sample(const float* mraw, const float* sraw, const float* eps,
                    float lo, float hi, float* action, float* mean, float* std,
                    int64_t n, dd_stream_t stream);
/* normalised Normal entropy per action dim (A/agent.py:361-367). */
int dd_actor_ent(const float* std, float lo, float hi, float* ent, int64_t n,
                 dd_stream_t stream);
/* per-row actor loss (A/agent.py:351-381, 'backprop'); adv = (mean, scale). */
int dd_actor_loss(const float* ret, const float* value, const float* adv,
                  const float* ent, const float* scale, const float* weight,
                  float* out, int64_t M, int A, dd_stream_t stream);
int dd_actor_bwd(const float* dA, const float* eps, const float* mean,
                 const float* std, const float* weight, const float* scale,
                 float coef, float lo, float hi, float* dmraw, float* dsraw,
                 int64_t M, int A, dd_stream_t stream);
/* discrete 'onehot' actor with the 'reinforce' gradient (A/nets.py:480-491,
 * A/agent.py:357-358,372-379): ent[m] = H(softmax(logit))/ln(A);
 * loss_row = (-logp(action)*score - scale*ent)*w, score = (ret-value-adv[0])*adv[1];
 * dlogit = coef * d loss_row / d logit (score and w are constants). */
int dd_actor_disc_ent(const float* logit, float* ent, int64_t M, int A,
                      dd_stream_t stream);
int dd_actor_disc_loss(const float* logit, const float* action, const float* ret,
                       const float* value, const float* adv, const float* ent,
                       const float* scale, const float* weight, float coef,
                       float* loss_row, float* dlogit, int64_t M, int A,
                       dd_stream_t stream);
/* lambda-return 'gve' + cont/weight (A/agent.py:256-259, :422-442). */
int dd_lambda_fwd(const float* rho, const float* vraw, const float* kraw,
                  const float* cont0, float gamma, float lam, float* reward,
                  float* value, float* cont, float* weight, float* ret, int H,
                  int N, dd_stream_t stream);
int dd_lambda_bwd(const float* rho, const float* vraw, const float* value,
                  const float* cont, const float* weight, const float* ret,
                  const float* adv, float coef, float gamma, float lam,
                  float* drho, float* dvraw, float* dkraw, int H, int N,
                  dd_stream_t stream);

/* ---- reductions, controllers, optimizer ------------------------------- */
/* rows of 8 doubles: [sum, sumsq, min, max, abssum, absmax, n, 0] */
int dd_stats_init(double* rows, int nrows, dd_stream_t stream);
int dd_stats(const float* X, int ldx, int rows, int cols, double* out,
             dd_stream_t stream);                    /* accumulates into out */
/* AutoAdapt 'mult' (A/tfutils.py:460-474); avg = sums[k]/denom. */
int dd_autoadapt(float* scale, const float* sums, int k, float denom,
                 float target, float lo, float hi, float vel, float thres,
                 int inverse, dd_stream_t stream);
/* Normalize.update + transform constants (A/tfutils.py:502-517), float64. */
int dd_normalize_update(double* state, const double* st, double decay,
                        double maxv, float* out, dd_stream_t stream);
int dd_sumsq(const float* g, int64_t n, double* out, dd_stream_t stream);
/* counters[ci]+=1; clip by global norm; weight decay on [0,n_decay); Adam
 * (A/tfutils.py:243-302). */
int dd_adam(float* p, const float* g, float* m, float* v, int64_t n,
            int64_t n_decay, float lr, double b1, double b2, float eps, float wd,
            float clip, float gscale, const double* sumsq, int64_t* counters,
            int ci, dd_stream_t stream);
/* VFunction.update_slow hard copy (A/agent.py:444-454). */
int dd_slow_sync(const float* src, float* dst, int64_t n, int64_t* counters,
                 int ci, int every, dd_stream_t stream);
int dd_counter_inc(int64_t* counters, int ci, dd_stream_t stream);

/* ---- noise (Philox4x32-10; stream layout in oracle/philox.py) --------- */
int dd_fill_gumbel(float* buf, int64_t n, uint64_t seed, uint32_t stream_id,
                   const int64_t* counters, int ci, dd_stream_t stream);
int dd_fill_normal(float* buf, int64_t n, uint64_t seed, uint32_t stream_id,
                   const int64_t* counters, int ci, dd_stream_t stream);

/* ---- layer machine: chains of GEMM + full-row epilogue layers in one persistent kernel ----
 * (csrc/dd_machine.cu).  One thread-block cluster (8 CTAs) owns a 128-row tile and runs every
 * layer of the program on it: TMA-fed tcgen05 3xTF32 tiles, LayerNorm row reductions over
 * distributed shared memory, cluster-local synchronisation between layers.  Replaces the
 * Linear -> LayerNorm -> elu stacks of A/nets.py:408-416, :566-602 (forward) and their reverse
 * chains.  Every GEMM operand is a pair (x, x_lo): x_lo = tf32(x - top19bits(x)), produced by
 * the layer that wrote x (out_lo), by dd_split_lo, or by dd_transpose_lo for weights.
 *   layer:  acc[M,N] = sum_seg  A_seg[M,k] . B[wn0 : wn0+N, wk0_seg : wk0_seg+k]^T
 *           B is a row-major [wrows, wcols] matrix (row stride ldw): the TRANSPOSED kernel for a
 *           forward Linear (x @ W), the kernel itself for an input gradient (dY @ W^T)
 *   DD_EPI_PLAIN   out = acc (+ add) (+ bias)
 *   DD_EPI_LN      z = acc (+ add) (+ bias); (mean, rstd) -> st; out = act(LN(z) * gamma + beta)
 *   DD_EPI_LNBWD   dh = acc (+ add); out = dz of  h = act(LN(z) * gamma + beta)  given the saved
 *                  z and st; dgamma / dbeta (+=) optional
 * a_lo == NULL marks an operand that is exact in tf32 (one-hot latents): its lo terms are skipped. */
#define DD_EPI_PLAIN 0
#define DD_EPI_LN 1
#define DD_EPI_LNBWD 2
/*   DD_EPI_GRU     N = 3 D: y = LN(acc) * gamma + beta over all 3 D columns (z, st, x1 = y saved);
 *                  out = GRU update of x0 = previous state (A/nets.py:149-160)
 *   DD_EPI_STATS   z = acc + bias; out = unimix logits per group of ipar classes (f0 = mix); with
 *                  x0 = Gumbel noise: x1 = one-hot of argmax(logit + noise) (A/nets.py:162-171,
 *                  A/tfutils.py:368-382)
 *   DD_EPI_ACTOR   N = 2 ipar columns [mean raw | std raw] (+ bias): x1 = mean = tanh(.), x2 = std =
 *                  (f1 - f0) sigmoid(.) + f0, out = mean + std * x0 (A/nets.py:461-468) */
/*   DD_EPI_STATSBWD input = d(one-hot sample) through the straight-through estimator, x0 = the
 *                  unimix logits of the forward pass: out = d(raw logits)
 *   DD_EPI_GRUBWD  N = D: input = d(new deter); x0 = previous deter, x1 = y, z / st of the forward GRU
 *                  layer: out [M, 3 D] = d(pre-norm gates), x2 [M, D] += d(previous deter) */
#define DD_EPI_GRU 3
#define DD_EPI_STATS 4
#define DD_EPI_ACTOR 5
#define DD_EPI_STATSBWD 6
#define DD_EPI_GRUBWD 7
#define DD_MAX_TRANSPOSE_JOBS 64
typedef struct DdSeg {
  const float* a;
  const float* a_lo;
  int lda, k, wk0;
  int lda_lo;       /* row stride of a_lo */
} DdSeg;
typedef struct DdLayer {
  int nseg;
  DdSeg seg[2];
  const float* w;
  const float* w_lo;
  int ldw, wrows, wcols, wn0;
  int N, epi, act;
  const float* add;
  int ld_add;
  const float* bias;
  const float* gamma;
  const float* beta;
  float* z;
  int ld_z;
  float* st;
  float* out;
  float* out_lo;
  int ld_out;
  int ld_lo;        /* row stride of out_lo */
  float* dgamma;
  float* dbeta;
  float* x[3];      /* epilogue-specific extra operands, see DD_EPI_* */
  int ldx[3];
  float f[2];
  int ipar;
} DdLayer;
/* bytes of device memory a compiled program of `nlayers` layers occupies */
size_t dd_program_bytes(int nlayers);
/* encodes the TMA descriptors and copies the program into table_dev (synchronous; call once,
 * outside stream capture).  rows = M of every layer. */
int dd_program_build(const DdLayer* layers, int nlayers, int rows, void* table_dev);
/* Host only (no device work): for every A segment the number of leading layers of the program that
 * write it, i.e. that must be complete before its first k-block is fetched; deps[2 * l + g], 0 for a
 * program input.  dd_program_build uses the same analysis to let early operands of a layer stream
 * while the previous layer's epilogue is still running. */
int dd_program_deps(const DdLayer* layers, int nlayers, int rows, int* deps);
int dd_program_run(const void* table_dev, int nlayers, int rows, dd_stream_t stream);
/* geometry of the machine: 1 = clusters of 8 CTAs (one per column slice), 2 = split-K clusters of
 * 16 (two CTAs per slice share its k-blocks; default when the device hosts such clusters) */
int dd_set_machine_ks(int ks);
/* after a failed run: 1 and the 8 diagnostic words of the stuck wait, else 0 */
int dd_program_debug(unsigned* out8);
/* profiling aids: switches that drop parts of the machine's pipeline (garbage results) and
 * per-CTA wait-cycle totals [cta][16] (index = wait code; 0 = kernel cycles) */
int dd_set_machine_debug(int flags);
int dd_machine_prof(int on, unsigned long long* out, int n);
/* lo[r, c] = tf32(x - top19bits(x)) for a row-strided matrix */
int dd_split_lo(const float* x, int ldx, float* lo, int ldlo, int rows, int cols,
                dd_stream_t stream);
/* for each job j over the row-major [rows_j, cols_j] matrix at p + src_j:
 *   mode 0: pt[dst_j + c*ldd_j + r] = p[src_j + r*cols_j + c]   (transposed copy, row stride ldd_j)
 *   mode 1: pt[dst_j + r*ldd_j + c] = p[src_j + r*cols_j + c]   (re-strided copy)
 * and the tf32 remainder of each element into pt_lo at the same index.  This is how the
 * optimizer's flat parameter buffer becomes the (w, w_lo) B operands of dd_program_build. */
int dd_transpose_lo(const float* p, float* pt, float* pt_lo, int njobs, const int* src,
                    const int* rows, const int* cols, const int* dst, const int* ldd,
                    const int* mode, dd_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif  /* DD_KERNELS_H_ */
