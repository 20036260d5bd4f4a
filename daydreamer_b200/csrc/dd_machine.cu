// The "layer machine": chains of  Y = epilogue(X . W)  layers executed by ONE persistent
// kernel, one thread-block cluster per 128-row tile.
//
// Why: the imagination rollout (A/agent.py:234-261), every MLP trunk (A/nets.py:408-416) and
// their reverse passes are chains of dependent [rows x K] . [K x N] products with a LayerNorm
// (over the FULL row) between them, and each 128-row tile of the batch runs through such a
// chain independently of every other tile.  So a cluster of CL = 8 CTAs owns a row tile: CTA r
// computes column slice r of every layer (tcgen05 kind::tf32 tiles 128 x bn, accumulators in
// TMEM), the 8 CTAs exchange the per-row LayerNorm partials through distributed shared memory,
// and the only synchronisation between consecutive layers is cluster-local -- no grid barrier, no
// kernel boundary.  Nothing on that path is a cluster-scope release/acquire: those compile to
// MEMBAR.ALL.GPU / CCTL.IVALL on sm_100a (measured: 10x the layer time).  Row partials travel as
// st.async ... mbarrier::complete_tx (the bytes ARE the signal), "my activations are written" is a
// relaxed remote add on a monotonic counter after ONE fence.acq_rel.cluster per layer per CTA.
//
// Operands reach the tensor core straight from TMA (cp.async.bulk.tensor, SWIZZLE_128B K-major
// tiles; SASS: UTMALDG): fp32-accurate 3xTF32 needs x = hi + lo with both halves in shared
// memory, so every operand exists in global memory as a pair (x, x_lo): the tensor core reads
// the top 19 bits of x itself as `hi` and x_lo = tf32(x - trunc19(x)) is written once by the
// producer of x (the previous layer's epilogue, or the optimizer for weights).  The A tile (the
// activations, identical for the 8 column-slice CTAs) is fetched ONCE per cluster: each CTA
// loads a 16-row slice and multicasts it to all 8; weights (B) are private column slices.
//
// Warp roles (Geo<KS>; KS = 1: clusters of 8, 352 threads; KS = 2: split-K clusters of 16, 224):
//   warp 0    weight producer (one lane issues): ring of 3 stages {A_hi, A_lo, B_hi, B_lo}; the
//             (w, w_lo) planes of a k-block are ONE 3-D box, so one TMA instruction per k-block
//   warp 6    activation producer: waits for the producing layer's `ready` count, then multicasts
//             its 16-row slice of A_hi / A_lo to the cluster (TMA issue costs ~250 cycles per
//             instruction, so A and B are issued by different warps)
//   warp 1    TMEM allocator + MMA issuer (one lane): per k-step  A_hi x [B_hi; B_lo]  (N = 2 bn)
//             and  A_lo x B_hi  (N = bn) -- two instructions for the three tf32 products --
//             then tcgen05.commit multicast to the `empty` barrier of every CTA in the cluster
//   others    epilogue warps (8 for KS = 1: two per TMEM lane quarter, each takes half the
//             columns): tcgen05.ld, bias / addend, the epilogue proper (LN forward / backward, GRU
//             cell and its reverse, latent statistics + Gumbel sample, actor head) with the
//             cross-CTA row reduction over DSMEM, and x / x_lo for the next layer written through
//             a per-warp transposing staging buffer (coalesced 128-byte rows)
// Scheduling across layers: the activation producer waits per A SEGMENT for the last layer that writes
// it (dd_program_build finds that from the pointers), so early operands of layer l stream during the
// epilogue of layer l - 1; a cluster with several row tiles runs them in interleaved pairs (dual_mode),
// each chain with its own `ready` counter, so one tile's k-loop fills the other's epilogue.
// A 2-second watchdog in every wait writes (role, barrier, layer) to mapped host memory and traps,
// so a protocol bug is a diagnosable error instead of a hung GPU.
#include <cuda.h>
#include <stdlib.h>
#include <string.h>

#include "dd_common.cuh"

namespace dd {
namespace lm {

constexpr int NS = 8;                 // column slices of a row tile (= activation row slices)
constexpr int BM = 128;               // rows per tile (= TMEM lanes)
constexpr int BK = 32;                // k-block: 32 fp32 = one 128-byte swizzle row
constexpr int BNMAX = 64;             // widest per-CTA column chunk
constexpr int A_BYTES = BM * BK * 4;                       // 16 KB
constexpr int B_BYTES = BNMAX * BK * 4;                    // 8 KB
constexpr int STAGE_BYTES = 2 * A_BYTES + 2 * B_BYTES;     // 48 KB: A_hi | A_lo | B_hi | B_lo
constexpr int SLICE_ROWS = BM / NS;                        // 16 rows of A loaded by each CTA
constexpr int SLICE_BYTES = SLICE_ROWS * BK * 4;           // 2 KB = two 8-row swizzle groups
constexpr int STG_LD = 36;
constexpr int STG_WARP_BYTES = 32 * STG_LD * 4;
constexpr int EP_BYTES = 208;
constexpr int TMEM_COLS = 256;        // 2 buffers x (main 64 | cross 64)

// byte-counted barriers (LayerNorm exchange, split-K reduction) come in PAIRS used alternately: a
// sender that runs one use ahead must not add its bytes to a phase the receiver has not completed
enum { BAR_FULL = 0, BAR_EMPTY = 32, BAR_ACCF = 64, BAR_ACCE = 80, BAR_STATS = 96, BAR_READY = 112,
       BAR_TMEM = 120, BAR_RED = 128, BAR_BYTES = 144 };

// device-side layer descriptor (built on the host by dd_program_build)
struct alignas(64) LmLayer {
  CUtensorMap a_hi[2], a_lo[2];   // per A segment: [rows, k] fp32, box {32, 16}, SWIZZLE_128B
  CUtensorMap w_hi, w_lo;         // B operand [n rows, k cols], box {32, bn}
  int nseg, seg_k[2], seg_wk0[2], seg_exact[2];
  int seg_dep[2];                 // leading layers of the program that must be complete before the segment's
                                  // first k-block (0: an external input); the last k-block waits for all of them
  int wn0;                        // first row of the B operand used by this layer
  int w3d;                        // w_hi is a 3-D map over the (w, w_lo) planes: one box per k-block
  int N, bn, nchunk, chunk_stride, nactive;
  int epi, act;
  int ld_add, ld_out, ld_z, ld_lo;
  const float* add;               // [M, N] added to the accumulator (beta = 1 / skip input)
  const float* bias;              // [N]
  const float* gamma;             // LayerNorm scale / bias
  const float* beta;
  float* z;                       // LN fwd: pre-norm output (saved for backward); LN bwd: saved pre-norm input
  float* st;                      // [M, 2] (mean, rstd): output of LN fwd, input of LN bwd
  float* out;                     // primary output  (x)
  float* out_lo;                  // its tf32 remainder (x_lo), or null
  float* dgamma;                  // LN bwd: parameter gradients (+=), or null
  float* dbeta;
  // epilogue-specific extras (GRU: x0 = previous deter, x1 = post-norm gates saved for backward;
  // STATS: x0 = Gumbel noise, x1 = one-hot sample; ACTOR: x0 = eps, x1 = mean, x2 = std)
  float* x[3];
  int ldx[3];
  float f[2];
  int ipar;
};

// ---- PTX helpers ---------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ uint32_t cluster_rank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ uint32_t mapa(uint32_t addr, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(rank));
  return r;
}
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_arrive_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes)
               : "memory");
}
// arrive on the barrier at the same offset in CTA `rank` of the cluster (release at cluster scope)
__device__ __forceinline__ void mbar_arrive_remote(uint32_t bar, uint32_t rank) {
  const uint32_t ra = mapa(bar, rank);
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(ra) : "memory");
}

// `ready` is a monotonically increasing counter, not an mbarrier: a producer that skips the wait of
// a first layer may be several layers away from the epilogues, which a one-bit phase cannot express
// (ptxas turns every .release.cluster / .acquire.cluster into MEMBAR.ALL.GPU / CCTL.IVALL, so the
// signal is ONE fence.acq_rel.cluster by one thread per CTA followed by relaxed remote adds)
__device__ __forceinline__ void count_add_remote(uint32_t addr, uint32_t rank) {
  const uint32_t ra = mapa(addr, rank);
  asm volatile("red.relaxed.cluster.shared::cluster.add.u32 [%0], %1;" ::"r"(ra), "r"(1u) : "memory");
}

// Diagnostics of a stuck wait: a wait that does not complete within ~2 s writes what it was
// waiting for into `dbg` (mapped host memory, survives the trap) and kills the kernel, so a
// protocol bug shows up as an error with a location instead of a hung GPU box.
struct Ctx {
  unsigned* dbg;
  int pair;                    // (tile, layer) pair index being processed
  unsigned long long* prof;    // optional: per-CTA wait-cycle totals by wait code (dd_set_machine_prof)
  int flags;                   // profiling switches (dd_set_machine_debug), results are garbage
  unsigned long long* trace;   // optional (CTA 0 only): clock64 at pipeline events, [role][event]
};
constexpr int TRACE_N = 2048;    // events per role
__device__ __forceinline__ void trace_ev(const Ctx& cx, int role, uint32_t idx) {
  if (cx.trace != nullptr && idx < (uint32_t)TRACE_N && (threadIdx.x & 31) == 0)
    cx.trace[role * TRACE_N + idx] = (unsigned long long)clock64();
}
enum { F_NOMMA = 1, F_NOA = 2, F_LOCALA = 4, F_NOB = 8, F_TMFENCE = 16, F_NORELAX = 32, F_NODUAL = 64 };
// Two row tiles per cluster, interleaved layer by layer ("dual"): when a cluster has several tiles to
// process, it runs them in pairs -- (layer l, tile A), (layer l, tile B), (layer l + 1, tile A), ... --
// so the k-loop of one tile runs during the epilogue of the other (a tile's own layers are dependent and
// leave the tensor pipe idle during every epilogue).  Every role walks the same sequence; each of the two
// chains has its own `ready` counter (a chain is strictly ordered, so its counter stays unambiguous).
__device__ __forceinline__ bool dual_mode(int ntiles, int nclusters, int flags) {
  return ntiles > nclusters && !(flags & F_NODUAL);
}
__device__ __noinline__ void stuck(const Ctx& cx, int code, uint32_t parity) {
  if (cx.dbg != nullptr) {
    cx.dbg[1] = (unsigned)code;
    cx.dbg[2] = blockIdx.x;
    cx.dbg[3] = threadIdx.x;
    cx.dbg[4] = parity;
    cx.dbg[5] = (unsigned)cx.pair;
    cx.dbg[0] = 0xDEADBEEFu;
    __threadfence_system();
  }
  __trap();
}
template <bool CLUSTER>
__device__ __forceinline__ bool mbar_try(uint32_t bar, uint32_t parity) {
  uint32_t done;
  if (CLUSTER)
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.b32 %0, 1, 0, p;\n\t}"
        : "=r"(done) : "r"(bar), "r"(parity) : "memory");
  else
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.b32 %0, 1, 0, p;\n\t}"
        : "=r"(done) : "r"(bar), "r"(parity) : "memory");
  return done != 0;
}
// the waiting path (watchdog, wait-cycle accounting) sits behind the first poll: a wait that is
// already satisfied -- the common case in the single-lane roles -- costs one poll and one branch
template <bool CLUSTER>
__device__ __forceinline__ void mbar_wait_slow(uint32_t bar, uint32_t parity, const Ctx& cx, int code) {
  unsigned long long t0 = 0;
  const long long c0 = cx.prof != nullptr ? clock64() : 0;
  for (uint32_t spin = 0;; ++spin) {
    if (mbar_try<CLUSTER>(bar, parity)) {
      if (cx.prof != nullptr && (threadIdx.x & 31) == 0)
        atomicAdd(cx.prof + code, (unsigned long long)(clock64() - c0));
      return;
    }
    if ((spin & 1023u) == 1023u) {
      unsigned long long t;
      asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
      if (t0 == 0) t0 = t;
      else if (t - t0 > 2000000000ull) stuck(cx, code, parity);
    }
  }
}
template <bool CLUSTER>
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity, const Ctx& cx, int code) {
  if (mbar_try<CLUSTER>(bar, parity)) return;
  mbar_wait_slow<CLUSTER>(bar, parity, cx, code);
}

__device__ __forceinline__ void count_wait(uint32_t addr, uint32_t target, const Ctx& cx, int code) {
  unsigned long long t0 = 0;
  const long long c0 = cx.prof != nullptr ? clock64() : 0;
  for (uint32_t spin = 0;; ++spin) {
    uint32_t v;
    asm volatile("ld.relaxed.cluster.shared::cta.u32 %0, [%1];" : "=r"(v) : "r"(addr) : "memory");
    if (v >= target) {
      asm volatile("fence.acq_rel.cluster;" ::: "memory");
      if (cx.prof != nullptr && (threadIdx.x & 31) == 0)
        atomicAdd(cx.prof + code, (unsigned long long)(clock64() - c0));
      return;
    }
    __nanosleep(32);
    if ((spin & 1023u) == 1023u) {
      unsigned long long t;
      asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
      if (t0 == 0) t0 = t;
      else if (t - t0 > 2000000000ull) stuck(cx, code, target);
    }
  }
}

__device__ __forceinline__ void tma_load(uint32_t dst, const CUtensorMap* map, int c0, int c1,
                                         uint32_t bar) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes"
      " [%0], [%1, {%3, %4}], [%2];"
      ::"r"(dst), "l"(reinterpret_cast<uint64_t>(map)), "r"(bar), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void tma_load_3d(uint32_t dst, const CUtensorMap* map, int c0, int c1, int c2,
                                            uint32_t bar) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes"
      " [%0], [%1, {%3, %4, %5}], [%2];"
      ::"r"(dst), "l"(reinterpret_cast<uint64_t>(map)), "r"(bar), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}
__device__ __forceinline__ void tma_load_mc(uint32_t dst, const CUtensorMap* map, int c0, int c1,
                                            uint32_t bar, uint16_t mask) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster"
      " [%0], [%1, {%4, %5}], [%2], %3;"
      ::"r"(dst), "l"(reinterpret_cast<uint64_t>(map)), "r"(bar), "h"(mask), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void tmap_acquire(const CUtensorMap* map, bool fence) {
  // The descriptor table is written once by a host copy that completes before any launch and
  // its memory is never reused for other descriptors (dd_program_build owns it), so the
  // descriptor cache cannot hold a stale copy; the explicit tensormap-proxy fence (several
  // microseconds per descriptor) is kept only as a debugging switch.
  if (fence)
    asm volatile("fence.proxy.tensormap::generic.acquire.gpu [%0], 128;"
                 ::"l"(reinterpret_cast<uint64_t>(map)) : "memory");
  asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(map)) : "memory");
}

// K-major SWIZZLE_128B operand tile (rows x 32 fp32, 128-byte rows, 8-row groups 1024 B apart):
// start >> 4 [0,14), LBO >> 4 [16,30) (=1, unused for swizzled K-major), SBO >> 4 [32,46),
// version 1 [46,48), layout SWIZZLE_128B = 2 [61,64)   (cute::UMMA::SmemDescriptor)
__device__ __forceinline__ uint64_t make_desc(uint32_t addr) {
  return (uint64_t)((addr & 0x3FFFFu) >> 4) | (1ull << 16) | ((uint64_t)(1024 >> 4) << 32) |
         (1ull << 46) | (2ull << 61);
}
// kind::tf32 instruction descriptor: D = F32 [4,6), A = B = TF32 (2) [7,10) [10,13), both K-major,
// N >> 3 [17,23), M >> 4 [24,29)   (cute::UMMA::InstrDescriptor)
__device__ __forceinline__ uint32_t make_idesc(int n) {
  return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);
}
__device__ __forceinline__ void mma_tf32(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc,
                                         uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void umma_commit_mc(uint32_t bar, uint16_t mask) {
  asm volatile(
      "tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
      ::"r"(bar), "h"(mask) : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];"
               ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tmem_ld32_nowait(uint32_t taddr, uint32_t (&v)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]),
        "=r"(v[7]), "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]),
        "=r"(v[14]), "=r"(v[15]), "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]),
        "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]),
        "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
      : "r"(taddr)
      : "memory");
}

// tf32 remainder of x: the tensor core uses the top 19 bits of x as `hi`; lo is rounded to tf32
// here so that the hardware's truncation of it is exact.
__device__ __forceinline__ float lo_of(float x) {
  const float hi = __uint_as_float(__float_as_uint(x) & 0xFFFFE000u);
  const float lo = x - hi;
  return __uint_as_float((__float_as_uint(lo) + 0x1000u) & 0xFFFFE000u);
}
// elu with the hardware exponential: |error| <= ~2e-7 absolute on (-1, 0], far inside the 1e-4 parity bar
__device__ __forceinline__ float elu_f(float y) { return y > 0.f ? y : __expf(y) - 1.f; }
// gate non-linearities on the hardware exponential / reciprocal: sigmoid to ~1e-6 relative, tanh to
// ~2e-7 absolute (r02 trace: the precise expf / tanhf / divide chains were 40 of the 65 K cycles of a
// GRU epilogue -- two warps per scheduler cannot hide their dependent-instruction latency).  The
// categorical head keeps the precise expf / logf: its argmax must reproduce the oracle's indices.
__device__ __forceinline__ float sigmoid_f(float x) { return __fdividef(1.f, 1.f + __expf(-x)); }
__device__ __forceinline__ float tanh_f(float x) { return 1.f - __fdividef(2.f, 1.f + __expf(2.f * x)); }

// ---- 32 x 32 sub-tile I/O through the warp's staging buffer ----------------------------------
// Thread `lane` owns row `lane` of the sub-tile in registers; global rows are `ld` floats apart.
// Vector path (ld % 4 == 0, 16-byte aligned, all 32 columns valid): each global access covers 4
// rows x 128 contiguous bytes.  Otherwise a guarded scalar path (narrow heads, N = 1 ... 31).
// r02 A/B (profiles/r02_machine_epilogue_io.txt): loads are fastest row-per-lane (the second 16 bytes
// of every 32-byte sector hit L1), stores are fastest through the staging buffer -- a row-per-lane
// store touches 32 lines per instruction (256 LSU wavefronts per sub-tile against ~100 staged) and
// made every forward epilogue 40 % longer.
#ifndef DD_LM_STAGED_STORE
#define DD_LM_STAGED_STORE 1
#endif
#ifndef DD_LM_STAGED_LOAD
#define DD_LM_STAGED_LOAD 0
#endif
__device__ __forceinline__ void sub_store(float* stg, const float (&v)[32], float* g, int ld,
                                          int rows_valid, int cols_valid, int lane) {
  const bool vec = cols_valid >= 32 && (ld & 3) == 0 && (reinterpret_cast<uintptr_t>(g) & 15) == 0;
  if (vec && !DD_LM_STAGED_STORE) {
    // every lane writes the 128 contiguous bytes of its own row: eight 16-byte stores, no trip
    // through shared memory and no warp barriers (the sectors are half-filled per instruction, and
    // the L2 write path is nowhere near a limit here)
    if (lane < rows_valid) {
      float4* row = reinterpret_cast<float4*>(g + (size_t)lane * ld);
#pragma unroll
      for (int j = 0; j < 8; ++j) row[j] = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
    }
  } else if (vec) {
    float4* srow = reinterpret_cast<float4*>(stg + lane * STG_LD);
#pragma unroll
    for (int j = 0; j < 8; ++j) srow[j] = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
    __syncwarp();
#pragma unroll
    for (int it = 0; it < 8; ++it) {
      const int r = it * 4 + (lane >> 3), c4 = (lane & 7) * 4;
      const float4 x = *reinterpret_cast<const float4*>(stg + r * STG_LD + c4);
      if (r < rows_valid) *reinterpret_cast<float4*>(g + (size_t)r * ld + c4) = x;
    }
    __syncwarp();
  } else if (lane < rows_valid) {
#pragma unroll
    for (int j = 0; j < 32; ++j)
      if (j < cols_valid) g[(size_t)lane * ld + j] = v[j];
  }
}
__device__ __forceinline__ void sub_load(float* stg, float (&v)[32], const float* g, int ld,
                                         int rows_valid, int cols_valid, int lane) {
  const bool vec = cols_valid >= 32 && (ld & 3) == 0 && (reinterpret_cast<uintptr_t>(g) & 15) == 0;
  if (vec && !DD_LM_STAGED_LOAD) {
    const float4* row = reinterpret_cast<const float4*>(g + (size_t)(lane < rows_valid ? lane : 0) * ld);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      float4 x = make_float4(0.f, 0.f, 0.f, 0.f);
      if (lane < rows_valid) x = row[j];
      v[4 * j] = x.x; v[4 * j + 1] = x.y; v[4 * j + 2] = x.z; v[4 * j + 3] = x.w;
    }
  } else if (vec) {
#pragma unroll
    for (int it = 0; it < 8; ++it) {
      const int r = it * 4 + (lane >> 3), c4 = (lane & 7) * 4;
      float4 x = make_float4(0.f, 0.f, 0.f, 0.f);
      if (r < rows_valid) x = *reinterpret_cast<const float4*>(g + (size_t)r * ld + c4);
      *reinterpret_cast<float4*>(stg + r * STG_LD + c4) = x;
    }
    __syncwarp();
    const float4* srow = reinterpret_cast<const float4*>(stg + lane * STG_LD);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const float4 x = srow[j];
      v[4 * j] = x.x; v[4 * j + 1] = x.y; v[4 * j + 2] = x.z; v[4 * j + 3] = x.w;
    }
    __syncwarp();
  } else {
#pragma unroll
    for (int j = 0; j < 32; ++j)
      v[j] = (lane < rows_valid && j < cols_valid) ? g[(size_t)lane * ld + j] : 0.f;
  }
}

// column sums over the warp's 32 rows: on return lane j holds sum_rows a_row[j] in a[0]
__device__ __forceinline__ void warp_colsum32(float (&a)[32], int lane) {
#pragma unroll
  for (int off = 16; off >= 1; off >>= 1) {
    const bool up = (lane & off) != 0;
#pragma unroll
    for (int i = 0; i < off; ++i) {
      const float send = up ? a[i] : a[i + off];
      const float keep = up ? a[i + off] : a[i];
      a[i] = keep + __shfl_xor_sync(0xffffffffu, send, off);
    }
  }
}

// ---- roles ------------------------------------------------------------------------------------
// Geometry.  A row tile is always cut into NS = 8 column slices; with KS = 2 the k-blocks of every
// layer are additionally split between two CTAs (cluster of 16: rank = kh * 8 + cs), so at ~1K rows
// (8 tiles) the whole GPU works on a layer instead of 64 SMs.  The two partial accumulators of a
// slice meet over distributed shared memory: each CTA of the pair sends the 32 columns it does not
// own and finishes the epilogue of the 32 columns it does -- the split halves the k-loop AND the
// epilogue of every CTA.
template <int KS>
struct Geo {
  static constexpr int CLS = NS * KS;
  // clusters of 8: EIGHT epilogue warps, two per TMEM lane quadrant, each finishing one 32-column
  // half of the CTA's 64-column slice (half the serial work per thread, twice the warps to hide the
  // latency of the tcgen05.ld / staging / global round trips); split-K: one half per CTA, 4 warps
  static constexpr int EW = KS == 1 ? 8 : 4;
  static constexpr int THREADS = 32 * (3 + EW);     // + weight producer, MMA issuer, activation producer
  static constexpr int STAGES = 3;
  static constexpr int SLOTS = CLS;                 // row partials exchanged per LayerNorm: one per owner CTA
  static constexpr int OFF_STATS = STAGES * STAGE_BYTES;            // float2 [2 parity][SLOTS][BM]
  static constexpr int STATS_BYTES = 2 * SLOTS * BM * 8;
  static constexpr int OFF_LOC = OFF_STATS + STATS_BYTES;           // float2 [BM]: the second warp group's row partials
  static constexpr int LOC_BYTES = EW == 8 ? BM * 8 : 0;
  static constexpr int OFF_RED = OFF_LOC + LOC_BYTES;               // float [2 parity][BM][32], swizzled
  static constexpr int RED_BYTES = KS == 1 ? 0 : 2 * BM * 32 * 4;
  static constexpr int OFF_STG = OFF_RED + RED_BYTES;               // per-warp 32 x 36 floats
  static constexpr int OFF_BAR = OFF_STG + EW * STG_WARP_BYTES;
  static constexpr int OFF_EP = OFF_BAR + BAR_BYTES;                // per-epilogue-warp layer view
  static constexpr int SMEM_BYTES = OFF_EP + EW * EP_BYTES;
  static_assert(SMEM_BYTES <= 232448, "shared memory budget");
};

struct Shared {
  char* smem;
  uint32_t base;      // shared-window address of smem
  uint32_t bars;      // of the barrier block
};

__device__ __forceinline__ int layer_kblocks(const LmLayer& L) {
  if (L.nseg == 0) return 0;
  int kb = (L.seg_k[0] + BK - 1) / BK;
  if (L.nseg > 1) kb += (L.seg_k[1] + BK - 1) / BK;
  return kb;
}
// k-blocks [lo, hi) of a chunk this CTA works on
// (only full-width layers, bn = 64, split their k range; a narrow layer runs on the kh = 0 CTAs)
template <int KS>
__device__ __forceinline__ void k_range(int kbc, int kh, int bn, int& lo, int& hi) {
  if (KS == 1) { lo = 0; hi = kbc; return; }
  if (bn != 64) { lo = 0; hi = kh == 0 ? kbc : 0; return; }
  const int mid = (kbc + 1) >> 1;
  lo = kh == 0 ? 0 : mid;
  hi = kh == 0 ? mid : kbc;
}

// Two producer warps share the stage ring: warp 0 streams this CTA's weight tiles (they never
// depend on the previous layer, so it runs ahead across layer boundaries on its own), warp 6
// multicasts this CTA's 16-row slice of the activations to the 8 CTAs that work on the same
// k-blocks (it is the one that waits for the previous layer's outputs).  Both arm the stage's
// `full` barrier with their own byte count.
// The producer and issuer roles run WARP-UNIFORM: all 32 lanes walk the loops with values the
// compiler can prove identical across the warp (shuffle-broadcast layer fields, counters), and only
// the TMA / MMA / commit instructions themselves sit under an elected-lane predicate.  Those
// instructions take their operands from uniform registers: when the loop ran on one lane of a
// diverged warp instead, every one of them was wrapped in an ELECT + 7 x R2UR + branch "waterfall"
// and a lone thread at ~5 cycles per dependent instruction needed 80 cycles per MMA and ~470
// cycles of ring bookkeeping per k-block (r02 trace), more than the tensor pipe needs for the
// MMAs themselves.  Stage index / phase are counters (no % or /), descriptors advance by constants.
struct Ring {
  uint32_t s, ph;                        // stage, phase parity of its `full` barrier
  __device__ __forceinline__ void next(int stages) {
    if (++s == (uint32_t)stages) { s = 0; ph ^= 1u; }
  }
};
__device__ __forceinline__ int ubc(int v) { return __shfl_sync(0xffffffffu, v, 0); }
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.b32 %0, 1, 0, p;\n\t}" : "=r"(pred));
  return pred != 0;
}
// the scalar fields of a layer the single-issue roles steer by, made warp-uniform
struct LayerU {
  int nseg, k0, k1, wk0, wk1, wn0, bn, nchunk, chunk_stride, nactive, w3d, exact0, exact1, dep0, dep1;
  __device__ __forceinline__ void load(const LmLayer& L) {
    nseg = ubc(L.nseg); k0 = ubc(L.seg_k[0]); k1 = ubc(L.seg_k[1]);
    wk0 = ubc(L.seg_wk0[0]); wk1 = ubc(L.seg_wk0[1]); wn0 = ubc(L.wn0); bn = ubc(L.bn);
    nchunk = ubc(L.nchunk); chunk_stride = ubc(L.chunk_stride); nactive = ubc(L.nactive);
    w3d = ubc(L.w3d); exact0 = ubc(L.seg_exact[0]); exact1 = ubc(L.seg_exact[1]);
    dep0 = ubc(L.seg_dep[0]); dep1 = ubc(L.seg_dep[1]);
  }
  __device__ __forceinline__ int kb0() const { return (k0 + BK - 1) / BK; }
  __device__ __forceinline__ int kbc() const { return kb0() + (nseg > 1 ? (k1 + BK - 1) / BK : 0); }
};

template <int KS>
__device__ void producer_b(const Shared& sh, const LmLayer* layers, int nlayers, int M, int rank,
                           int cluster_id, int nclusters, int lane, Ctx cx) {
  constexpr int STAGES = Geo<KS>::STAGES;
  const int ntiles = (M + BM - 1) / BM;
  const int cs = rank % NS, kh = rank / NS;
  Ring rg = {0u, 0u};
  uint32_t it = 0;
  int pair = 0;
  const bool dual = dual_mode(ntiles, nclusters, cx.flags);
  const int tstep = dual ? 2 * nclusters : nclusters;
  for (int t0 = cluster_id; t0 < ntiles; t0 += tstep) {
    const int ng = (dual && t0 + nclusters < ntiles) ? 2 : 1;      // tiles of this group
    for (int l = 0; l < nlayers; ++l) {
      const LmLayer& L = layers[l];
      LayerU u;
      u.load(L);
      if (u.nseg == 0) { pair += ng; continue; }
      if (elect_one()) {
        tmap_acquire(&L.w_hi, (cx.flags & F_TMFENCE) != 0);
        if (!u.w3d) tmap_acquire(&L.w_lo, (cx.flags & F_TMFENCE) != 0);
      }
      const int kb0 = u.kb0(), kbc = u.kbc();
      int k_lo, k_hi;
      k_range<KS>(kbc, kh, u.bn, k_lo, k_hi);
      const int bn = u.bn;
      const int wk0 = u.wk0, wk1 = u.wk1 - kb0 * BK;     // k = wk + r * BK
      const bool active = cs < u.nactive && !(cx.flags & F_NOB);
      const uint32_t bytes = active ? 2u * (uint32_t)bn * BK * 4 : 0u;
      const CUtensorMap* whi = &L.w_hi;
      const CUtensorMap* wlo = &L.w_lo;
      for (int sub = 0; sub < ng; ++sub, ++pair) {       // the same weight tiles for both row tiles
        cx.pair = pair;
        int n = u.wn0 + cs * bn;
        for (int c = 0; c < u.nchunk; ++c, n += u.chunk_stride) {
          for (int r = k_lo; r < k_hi; ++r, ++it) {
            const uint32_t full = sh.bars + BAR_FULL + 8 * rg.s;
            mbar_wait<false>(sh.bars + BAR_EMPTY + 8 * rg.s, rg.ph ^ 1u, cx, 1);   // smem reuse only
            trace_ev(cx, 0, 4 * it);
            if (elect_one()) {
              mbar_arrive_tx(full, bytes);
              if (active) {
                const uint32_t st = sh.base + rg.s * STAGE_BYTES + 2 * A_BYTES;
                const int k = (r < kb0 ? wk0 : wk1) + r * BK;
                if (u.w3d) {               // rows [0, bn) = w, rows [bn, 2 bn) = w_lo: one box
                  tma_load_3d(st, whi, k, n, 0, full);
                } else {
                  tma_load(st, whi, k, n, full);
                  tma_load(st + (uint32_t)bn * (BK * 4), wlo, k, n, full);
                }
              }
            }
            trace_ev(cx, 0, 4 * it + 1);
            rg.next(STAGES);
          }
        }
      }
    }
  }
}

template <int KS>
__device__ void producer_a(const Shared& sh, const LmLayer* layers, int nlayers, int M, int rank,
                           int cluster_id, int nclusters, int lane, Ctx cx) {
  constexpr int STAGES = Geo<KS>::STAGES;
  const int ntiles = (M + BM - 1) / BM;
  const int cs = rank % NS, kh = rank / NS;
  const uint16_t mask = (uint16_t)(0xFFu << (NS * kh));      // the CTAs sharing my k-blocks
  const bool noa = (cx.flags & F_NOA) != 0;
  const int relax = (cx.flags & F_NORELAX) ? 0 : 1;
  Ring rg = {0u, 0u};
  uint32_t it = 0;
  int pair = 0;
  int have0 = 0, have1 = 0;                // layers of chain 0 / 1 known complete on every CTA
  int base0 = 0, base1 = 0;                // layers the chains completed in earlier tile groups
  const bool dual = dual_mode(ntiles, nclusters, cx.flags);
  const int tstep = dual ? 2 * nclusters : nclusters;
  for (int t0 = cluster_id; t0 < ntiles; t0 += tstep) {
    const int ng = (dual && t0 + nclusters < ntiles) ? 2 : 1;
    for (int l = 0; l < nlayers; ++l) {
      const LmLayer& L = layers[l];
      LayerU u;
      u.load(L);
      if (u.nseg == 0) { pair += ng; continue; }
      if (elect_one()) {
        for (int g = 0; g < u.nseg; ++g) {
          tmap_acquire(&L.a_hi[g], (cx.flags & F_TMFENCE) != 0);
          if (!(g ? u.exact1 : u.exact0)) tmap_acquire(&L.a_lo[g], (cx.flags & F_TMFENCE) != 0);
        }
      }
      const int kb0 = u.kb0(), kbc = u.kbc();
      int k_lo, k_hi;
      k_range<KS>(kbc, kh, u.bn, k_lo, k_hi);
      const bool exact0 = u.exact0 != 0, exact1 = u.exact1 != 0;
      const CUtensorMap* ah0 = &L.a_hi[0];
      const CUtensorMap* al0 = &L.a_lo[0];
      const CUtensorMap* ah1 = &L.a_hi[1];
      const CUtensorMap* al1 = &L.a_lo[1];
      // A segment is fetched as soon as the layers that WRITE it are complete on every CTA
      // (seg_dep, found by dd_program_build), not when the previous layer is: the k-blocks of an
      // early operand (the carried deter of the GRU, the latent of img_in / the actor) run while
      // the previous layer's epilogue is still busy.  The last k-block of the first chunk waits
      // for everything before this layer, so no accumulator of layer l is complete -- and no
      // epilogue of layer l starts anywhere -- before layer l - 1 is published, exactly as before.
      for (int sub = 0; sub < ng; ++sub, ++pair) {
        cx.pair = pair;
        const int m = (t0 + sub * nclusters) * BM + cs * SLICE_ROWS;
        const int base = sub ? base1 : base0;          // this chain's layers before the current tile
        const uint32_t ready = sh.bars + BAR_READY + 4 * sub;
        for (int c = 0; c < u.nchunk; ++c) {
          for (int r = k_lo; r < k_hi; ++r, ++it) {
            const bool s1 = r >= kb0;
            const bool exact = s1 ? exact1 : exact0;
            const uint32_t full = sh.bars + BAR_FULL + 8 * rg.s;
            const int rel = ((c == 0 && r == k_hi - 1) || relax == 0) ? l : (s1 ? u.dep1 : u.dep0);
            if (rel > 0 && base + rel > (sub ? have1 : have0)) {   // rel == 0: inputs of the program
              if (sub) have1 = base + rel; else have0 = base + rel;
              count_wait(ready, (uint32_t)(base + rel) * Geo<KS>::CLS, cx, 2);
              asm volatile("fence.proxy.async;" ::: "memory");
            }
            mbar_wait<false>(sh.bars + BAR_EMPTY + 8 * rg.s, rg.ph ^ 1u, cx, 9);
            trace_ev(cx, 0, 4 * it + 2);
            if (elect_one()) {
              mbar_arrive_tx(full, noa ? 0u : (exact ? 1u : 2u) * A_BYTES);
              if (!noa) {
                const uint32_t st = sh.base + rg.s * STAGE_BYTES + cs * SLICE_BYTES;
                const int k = (s1 ? r - kb0 : r) * BK;
                tma_load_mc(st, s1 ? ah1 : ah0, k, m, full, mask);
                if (!exact) tma_load_mc(st + A_BYTES, s1 ? al1 : al0, k, m, full, mask);
              }
            }
            trace_ev(cx, 0, 4 * it + 3);
            rg.next(STAGES);
          }
        }
      }
    }
    base0 += nlayers;
    if (ng == 2) base1 += nlayers;
  }
}

template <int KS>
__device__ void mma_issuer(const Shared& sh, const LmLayer* layers, int nlayers, int M, int rank,
                           int cluster_id, int nclusters, int lane, uint32_t tmem_base, Ctx cx) {
  constexpr int STAGES = Geo<KS>::STAGES;
  const int ntiles = (M + BM - 1) / BM;
  const int cs = rank % NS, kh = rank / NS;
  const uint16_t mask = (uint16_t)(0xFFu << (NS * kh));
  const bool nomma = (cx.flags & F_NOMMA) != 0;
  // descriptors of stage 0; stage s adds s * STAGE_BYTES and k-step ks adds 32 bytes, both in the
  // descriptor's 16-byte units (the address field cannot carry: every sum is a valid smem address)
  const uint64_t d_ahi0 = make_desc(sh.base), d_alo0 = make_desc(sh.base + A_BYTES);
  const uint64_t d_b0 = make_desc(sh.base + 2 * A_BYTES);
  Ring rg = {0u, 0u};
  uint32_t it = 0, ac = 0;
  int pair = 0;
  const bool dual = dual_mode(ntiles, nclusters, cx.flags);
  const int tstep = dual ? 2 * nclusters : nclusters;
  for (int t0 = cluster_id; t0 < ntiles; t0 += tstep) {
    const int ng = (dual && t0 + nclusters < ntiles) ? 2 : 1;
    for (int l = 0; l < nlayers; ++l) {
      LayerU u;
      u.load(layers[l]);
      if (u.nseg == 0) { pair += ng; continue; }   // epilogue-only layer
      const int kb0 = u.kb0(), kbc = u.kbc();
      int k_lo, k_hi;
      k_range<KS>(kbc, kh, u.bn, k_lo, k_hi);
      if (k_hi <= k_lo) { pair += ng; continue; }  // nothing of this layer's k range is mine
      const bool active = cs < u.nactive;
      const int bn = u.bn;
      const uint32_t idesc = make_idesc(bn), idesc2 = make_idesc(2 * bn);
      // k-steps of the last (possibly partial) k-block of each segment
      const int last0 = kb0 - 1, last1 = kbc - 1;
      const int ks_last0 = ((u.k0 - last0 * BK) + 7) >> 3;
      const int ks_last1 = u.nseg > 1 ? ((u.k1 - (last1 - kb0) * BK) + 7) >> 3 : 4;
      const bool exact0 = u.exact0 != 0, exact1 = u.exact1 != 0;
      const bool issue = active && !nomma;
      for (int sub = 0; sub < ng; ++sub, ++pair) {
      cx.pair = pair;
      for (int c = 0; c < u.nchunk; ++c) {
        const uint32_t buf = ac & 1;
        if (active) {
          mbar_wait<false>(sh.bars + BAR_ACCE + 8 * buf, ((ac >> 1) & 1) ^ 1, cx, 3);
          asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        }
        // accumulator buffer: columns [0, bn) main (hi.hi), [bn, 2 bn) the small cross terms
        const uint32_t d_main = tmem_base + buf * 128, d_cross = d_main + (uint32_t)bn;
        uint32_t acc = 0;
        for (int r = k_lo; r < k_hi; ++r, ++it) {
          const bool s1 = r >= kb0;
          const bool exact = s1 ? exact1 : exact0;
          const int ksteps = r == last0 ? ks_last0 : (r == last1 ? ks_last1 : 4);
          mbar_wait<false>(sh.bars + BAR_FULL + 8 * rg.s, rg.ph, cx, 4);
          asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
          trace_ev(cx, 1, 4 * it);
          if (elect_one()) {
            if (issue) {
              const uint64_t so = (uint64_t)(rg.s * (STAGE_BYTES >> 4));
              const uint64_t dah = d_ahi0 + so, dal = d_alo0 + so, dbh = d_b0 + so;
              // B_hi and B_lo are rows [0, bn) and [bn, 2 bn) of ONE K-major tile: a single
              // N = 2 bn product gives hi.hi (main) and hi.lo (cross) with one read of A_hi
              if (ksteps == 4 && !exact) {
                mma_tf32(d_main, dah, dbh, idesc2, acc);
                mma_tf32(d_cross, dal, dbh, idesc, 1u);
#pragma unroll
                for (int ks = 1; ks < 4; ++ks) {
                  mma_tf32(d_main, dah + 2 * ks, dbh + 2 * ks, idesc2, 1u);
                  mma_tf32(d_cross, dal + 2 * ks, dbh + 2 * ks, idesc, 1u);
                }
              } else if (ksteps == 4) {
                mma_tf32(d_main, dah, dbh, idesc2, acc);
#pragma unroll
                for (int ks = 1; ks < 4; ++ks) mma_tf32(d_main, dah + 2 * ks, dbh + 2 * ks, idesc2, 1u);
              } else {
                for (int ks = 0; ks < ksteps; ++ks) {
                  mma_tf32(d_main, dah + 2 * ks, dbh + 2 * ks, idesc2, ks == 0 ? acc : 1u);
                  if (!exact) mma_tf32(d_cross, dal + 2 * ks, dbh + 2 * ks, idesc, 1u);
                }
              }
            }
            // this stage may be refilled once every CTA that received it has read it
            umma_commit_mc(sh.bars + BAR_EMPTY + 8 * rg.s, mask);
            if (active && r == k_hi - 1) umma_commit(sh.bars + BAR_ACCF + 8 * buf);
          }
          acc = 1;
          trace_ev(cx, 1, 4 * it + 1);
          rg.next(STAGES);
        }
        if (active) ++ac;
      }
      }
    }
  }
}

// ---- epilogue ----------------------------------------------------------------------------------
// the epilogue's view of a layer, copied into shared memory once per (tile, layer) and warp: the
// descriptor itself lives in global memory and would be re-read after every global store
struct EP {
  int nseg, N, bn, nchunk, chunk_stride, nactive, epi, act, ld_add, ld_out, ld_z, ld_lo;
  const float *add, *bias, *gamma, *beta;
  float *z, *st, *out, *out_lo, *dgamma, *dbeta;
  float* x[3];
  int ldx[3];
  float f[2];
  int ipar;
  int kbc;            // k-blocks per chunk (0: epilogue-only layer)
};
static_assert(sizeof(EP) <= EP_BYTES, "EP slot");
__device__ __forceinline__ EP load_ep(const LmLayer& L) {
  EP e;
  e.nseg = L.nseg; e.N = L.N; e.bn = L.bn; e.nchunk = L.nchunk; e.chunk_stride = L.chunk_stride;
  e.nactive = L.nactive; e.epi = L.epi; e.act = L.act;
  e.ld_add = L.ld_add; e.ld_out = L.ld_out; e.ld_z = L.ld_z; e.ld_lo = L.ld_lo;
  e.add = L.add; e.bias = L.bias; e.gamma = L.gamma; e.beta = L.beta;
  e.z = L.z; e.st = L.st; e.out = L.out; e.out_lo = L.out_lo; e.dgamma = L.dgamma; e.dbeta = L.dbeta;
  for (int i = 0; i < 3; ++i) { e.x[i] = L.x[i]; e.ldx[i] = L.ldx[i]; }
  e.f[0] = L.f[0]; e.f[1] = L.f[1]; e.ipar = L.ipar;
  e.kbc = layer_kblocks(L);
  return e;
}
// 32 consecutive per-column parameters (16-byte aligned: offsets are multiples of 32 floats)
__device__ __forceinline__ void load32(const float* p, float (&o)[32]) {
  if ((reinterpret_cast<uintptr_t>(p) & 15) == 0) {
#pragma unroll
    for (int q = 0; q < 8; ++q) {
      const float4 x = reinterpret_cast<const float4*>(p)[q];
      o[4 * q] = x.x; o[4 * q + 1] = x.y; o[4 * q + 2] = x.z; o[4 * q + 3] = x.w;
    }
  } else {
#pragma unroll
    for (int j = 0; j < 32; ++j) o[j] = p[j];
  }
}

// per-thread epilogue state: the same for every layer of a tile
struct Epi {
  Shared sh;
  Ctx cx;
  uint32_t tmem_base;
  uint32_t ac;        // accumulator buffers consumed
  uint32_t xc;        // cross-CTA row exchanges done
  uint32_t rc;        // split-K reductions done
  int rank, cs, kh, warp, quad, lane;
  int ew, grp;        // epilogue warp index 0 .. EW-1; its 32-column half (0 unless EW == 8)
  float* stg;
  int wrow0, rows_valid, trow;
  uint32_t stats_off, red_off, loc_off;
};

// main + cross accumulator columns [cc*32, cc*32+32) of buffer `buf`, this thread's row
__device__ __forceinline__ void acc_load32(uint32_t tmem_base, uint32_t buf, int quad, int cc, int bn,
                                           float (&v)[32]) {
  uint32_t a[32], b[32];
  const uint32_t t = tmem_base + ((uint32_t)(quad * 32) << 16) + buf * 128 + cc * 32;
  tmem_ld32_nowait(t, a);
  tmem_ld32_nowait(t + (uint32_t)bn, b);
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(a[j]) + __uint_as_float(b[j]);
}

// Does this CTA finish (own) columns of the layer, how many 32-column sub-tiles per chunk, and
// where sub-tile h of chunk c starts.  KS = 1: both halves of the slice; KS = 2: one half each
// (bn = 64) or the whole slice on the kh = 0 CTA (bn = 32).
template <int KS>
__device__ __forceinline__ bool owns(const EP& L, const Epi& e) {
  if (e.cs >= L.nactive) return false;
  return KS == 1 || L.bn == 64 || e.kh == 0;
}
template <int KS>
__device__ __forceinline__ int nsub(const EP& L) { return KS == 1 ? (L.bn >> 5) : 1; }
template <int KS>
__device__ __forceinline__ int sub_col(const EP& L, const Epi& e, int c, int h) {
  const int base = c * L.chunk_stride + e.cs * L.bn;
  if (KS == 1) return base + h * 32;
  return base + (L.bn == 64 ? e.kh * 32 : 0);
}
// with eight epilogue warps a thread finishes only the half of its warp group
template <int KS>
__device__ __forceinline__ bool mine(const Epi& e, int h) { return Geo<KS>::EW == 4 || h == e.grp; }
// The LayerNorm partials of a row come from `n_slots` threads (one per owned 32-column sub-tile per
// chunk set): 8 CTAs x 2 halves, or 16 split-K CTAs x 1, or 8 x 1 for 32-wide slices.  Slot of this
// thread, and the CTAs that own columns (ranks 0 .. n_owners-1).
template <int KS>
__device__ __forceinline__ int n_owners(const EP& L) { return (KS == 2 && L.bn == 64) ? 2 * NS : NS; }
template <int KS>
__device__ __forceinline__ int n_slots(const EP& L) { return n_owners<KS>(L); }
template <int KS>
__device__ __forceinline__ int my_slot(const EP& L, const Epi& e) {
  return (KS == 2 && L.bn == 64) ? e.rank : e.cs;
}

__device__ __forceinline__ bool has_k(const EP& L, const Epi& e, int ks) {
  if (L.kbc == 0) return false;
  if (ks == 1) return true;
  if (L.bn != 64) return e.kh == 0;
  const int mid = (L.kbc + 1) >> 1;
  return e.kh == 0 ? mid > 0 : L.kbc > mid;
}

__device__ __forceinline__ void acc_wait(Epi& e) {
  mbar_wait<false>(e.sh.bars + BAR_ACCF + 8 * (e.ac & 1), (e.ac >> 1) & 1, e.cx, 6);
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  if (e.warp == 2) trace_ev(e.cx, 2, 16 * e.cx.pair + 1);
}
__device__ __forceinline__ void acc_release(Epi& e) {
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncwarp();
  if (e.lane == 0) mbar_arrive(e.sh.bars + BAR_ACCE + 8 * (e.ac & 1));
  ++e.ac;
}

// Sub-tile h of the current chunk as a raw sum: the accumulator (main + cross) of this CTA's
// k-blocks, plus -- split-K -- the partner CTA's partial of the same columns, which arrives over
// distributed shared memory (st.async into a swizzled row buffer, bytes counted by `red`).
// Called by every epilogue thread of every active CTA for h = 0 .. nsub-1 of every chunk, in order.
template <int KS>
__device__ __forceinline__ void sub_fetch(const EP& L, Epi& e, int h, float (&v0)[32]) {
  const bool hk = has_k(L, e, KS);
  if (KS == 1) {           // this warp's group reads only half h of the accumulator
    if (hk) {
      acc_wait(e);
      acc_load32(e.tmem_base, e.ac & 1, e.quad, h, L.bn, v0);
      acc_release(e);
    } else {
#pragma unroll
      for (int j = 0; j < 32; ++j) v0[j] = 0.f;
    }
    return;
  }
  float v1[32];
  // KS == 2.  Narrow layers (bn = 32) are not split: the kh = 0 CTA has the whole sum.
  if (L.bn != 64) {
    if (hk) {
      acc_wait(e);
      acc_load32(e.tmem_base, e.ac & 1, e.quad, 0, L.bn, v0);
      acc_release(e);
    } else {
#pragma unroll
      for (int j = 0; j < 32; ++j) v0[j] = 0.f;
    }
    return;
  }
  // v0 = the 32 columns I keep; the 32 my partner finishes are read first and sent right away
  // (v1 is only a staging array), so both halves are never live together
  const uint32_t use = e.rc, par = use & 1;
  const uint32_t bar = e.sh.bars + BAR_RED + 8 * par;
  if (e.warp == 2 && e.lane == 0) mbar_arrive_tx(bar, BM * 32 * 4);
  if (hk) {
    acc_wait(e);
    acc_load32(e.tmem_base, e.ac & 1, e.quad, 1 - e.kh, L.bn, v1);
  } else {
#pragma unroll
    for (int j = 0; j < 32; ++j) v1[j] = 0.f;
  }
  {
    const uint32_t partner = (uint32_t)(e.rank ^ NS);
    const uint32_t row = e.sh.base + e.red_off + (par * BM + e.trow) * 128;
    const uint32_t rrow = mapa(row, partner), rbar = mapa(bar, partner);
#pragma unroll
    for (int q = 0; q < 8; ++q) {
      const uint32_t a = rrow + (uint32_t)((q ^ (e.trow & 7)) << 4);
      asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.v4.f32 [%0], {%1, %2, %3, %4}, [%5];"
                   ::"r"(a), "f"(v1[4 * q]), "f"(v1[4 * q + 1]), "f"(v1[4 * q + 2]), "f"(v1[4 * q + 3]), "r"(rbar)
                   : "memory");
    }
  }
  if (hk) {
    acc_load32(e.tmem_base, e.ac & 1, e.quad, e.kh, L.bn, v0);
    acc_release(e);
  } else {
#pragma unroll
    for (int j = 0; j < 32; ++j) v0[j] = 0.f;
  }
  mbar_wait<false>(bar, (use >> 1) & 1, e.cx, 10);
  {
    const char* row = e.sh.smem + e.red_off + (par * BM + e.trow) * 128;
#pragma unroll
    for (int q = 0; q < 8; ++q) {
      const float4 x = *reinterpret_cast<const float4*>(row + ((q ^ (e.trow & 7)) << 4));
      v0[4 * q] += x.x; v0[4 * q + 1] += x.y; v0[4 * q + 2] += x.z; v0[4 * q + 3] += x.w;
    }
  }
  ++e.rc;
}

// + addend + bias on an owned sub-tile
__device__ __forceinline__ void add_inputs(const EP& L, Epi& e, int cb, int cols_valid, float (&v)[32]) {
  if (L.add != nullptr) {
    float a[32];
    sub_load(e.stg, a, L.add + (size_t)e.wrow0 * L.ld_add + cb, L.ld_add, e.rows_valid, cols_valid, e.lane);
#pragma unroll
    for (int j = 0; j < 32; ++j) v[j] += a[j];
  }
  if (L.bias != nullptr) {
    if (cols_valid >= 32) {
      float b[32];
      load32(L.bias + cb, b);
#pragma unroll
      for (int j = 0; j < 32; ++j) v[j] += b[j];
    } else {
#pragma unroll
      for (int j = 0; j < 32; ++j)
        if (j < cols_valid) v[j] += L.bias[cb + j];
    }
  }
}

__device__ __forceinline__ void store_with_lo(const EP& L, Epi& e, float (&v)[32], int cb, int cols_valid) {
  sub_store(e.stg, v, L.out + (size_t)e.wrow0 * L.ld_out + cb, L.ld_out, e.rows_valid, cols_valid, e.lane);
  if (L.out_lo != nullptr) {
#pragma unroll
    for (int j = 0; j < 32; ++j) v[j] = lo_of(v[j]);
    sub_store(e.stg, v, L.out_lo + (size_t)e.wrow0 * L.ld_lo + cb, L.ld_lo, e.rows_valid, cols_valid, e.lane);
  }
}

// publish this thread's pair of row partials to every owner CTA, then collect all of them.
// The remote stores carry their own completion (st.async ... mbarrier::complete_tx): the destination's
// barrier counts the bytes, so no fence and no separate arrive is needed.
// `welford`: (a, b) = (mean, M2) over `cnt` columns, merged by Chan's formula; else plain sums.
// With two warp groups per row the second group's partial is first folded into the first group's
// through shared memory (a 64-thread named barrier per TMEM quadrant), so one thread per row sends.
template <int KS>
__device__ __forceinline__ void row_send(const EP& L, Epi& e, float a, float b, float& cnt, bool welford) {
  constexpr int SLOTS = Geo<KS>::SLOTS;
  const bool two = Geo<KS>::EW == 8 && L.bn == 64;
  if (two) {
    float2* loc = reinterpret_cast<float2*>(e.sh.smem + e.loc_off) + e.trow;
    if (e.grp == 1) *loc = make_float2(a, b);
    asm volatile("bar.sync %0, 64;" ::"r"(2 + e.quad) : "memory");
    if (e.grp == 0) {
      const float2 o = *loc;
      if (welford) {                 // equal counts: mean of means, M2 += cnt/2 * delta^2
        const float d = o.x - a;
        a += 0.5f * d;
        b += o.y + d * d * (cnt * 0.5f);
      } else {
        a += o.x;
        b += o.y;
      }
    }
    cnt *= 2.f;
  }
  const int nown = n_owners<KS>(L), nsl = n_slots<KS>(L), slot_i = my_slot<KS>(L, e);
  const uint32_t par = e.xc & 1;
  const uint32_t bar = e.sh.bars + BAR_STATS + 8 * par;
  if (e.warp == 2 && e.lane == 0) mbar_arrive_tx(bar, (uint32_t)nsl * BM * 8);   // bytes I will receive
  if (!two || e.grp == 0) {
    const uint32_t slot = e.sh.base + e.stats_off + ((par * SLOTS + slot_i) * BM + e.trow) * 8;
#pragma unroll
    for (int d = 0; d < Geo<KS>::CLS; ++d) {
      if (d < nown) {              // owners are ranks 0 .. nown-1 in both geometries
        const uint32_t ra = mapa(slot, d), rb = mapa(bar, d);
        asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.v2.f32 [%0], {%1, %2}, [%3];"
                     ::"r"(ra), "f"(a), "f"(b), "r"(rb) : "memory");
      }
    }
  }
}
// ... and the other half: wait for every owner's partials of my rows (work that does not need them
// -- storing the pre-norm values -- goes between row_send and row_recv)
template <int KS>
__device__ __forceinline__ void row_recv(const EP& L, Epi& e, float2 (&parts)[Geo<KS>::SLOTS]) {
  constexpr int SLOTS = Geo<KS>::SLOTS;
  const int nsl = n_slots<KS>(L);
  const uint32_t par = e.xc & 1;
  const uint32_t bar = e.sh.bars + BAR_STATS + 8 * par;
  mbar_wait<false>(bar, (e.xc >> 1) & 1, e.cx, 5);
  const float2* mine_ = reinterpret_cast<const float2*>(e.sh.smem + e.stats_off) + par * SLOTS * BM + e.trow;
#pragma unroll
  for (int d = 0; d < SLOTS; ++d)
    if (d < nsl) parts[d] = mine_[d * BM];
  ++e.xc;
}
template <int KS>
__device__ __forceinline__ void row_exchange(const EP& L, Epi& e, float a, float b, float& cnt, bool welford,
                                             float2 (&parts)[Geo<KS>::SLOTS]) {
  row_send<KS>(L, e, a, b, cnt, welford);
  row_recv<KS>(L, e, parts);
}

// running (count, mean, M2) of a row += one 32-column sub-tile
__device__ __forceinline__ void welford32(const float (&v)[32], float& mean, float& m2, float& cnt) {
  float s = 0.f;
#pragma unroll
  for (int j = 0; j < 32; ++j) s += v[j];
  const float mu = s * (1.f / 32.f);
  float q = 0.f;
#pragma unroll
  for (int j = 0; j < 32; ++j) { const float d = v[j] - mu; q += d * d; }
  const float delta = mu - mean, tot = cnt + 32.f;
  mean += delta * (32.f / tot);
  m2 += q + delta * delta * (cnt * 32.f / tot);
  cnt = tot;
}
template <int KS>
__device__ __forceinline__ void ln_recv(const EP& L, Epi& e, float cnt, float& gm, float& rstd);
// merge the owners' equal-count partials into (mean, rstd) of the full row
template <int KS>
__device__ __forceinline__ void ln_finish(const EP& L, Epi& e, float mean, float m2, float cnt,
                                          float& gm, float& rstd) {     // cnt by value: doubled by the local merge
  row_send<KS>(L, e, mean, m2, cnt, true);
  ln_recv<KS>(L, e, cnt, gm, rstd);
}
template <int KS>
__device__ __forceinline__ void ln_recv(const EP& L, Epi& e, float cnt, float& gm, float& rstd) {
  float2 parts[Geo<KS>::SLOTS];
  row_recv<KS>(L, e, parts);
  const int nsl = n_slots<KS>(L);
  gm = 0.f;
  for (int d = 0; d < nsl; ++d) gm += parts[d].x;
  gm /= (float)nsl;
  float gq = 0.f;
  for (int d = 0; d < nsl; ++d) { const float dl = parts[d].x - gm; gq += parts[d].y + cnt * dl * dl; }
  rstd = rsqrtf(gq / (float)L.N + DD_LN_EPS);
  if (L.st != nullptr && e.rank == 0 && e.grp == 0 && e.lane < e.rows_valid) {
    L.st[(size_t)(e.wrow0 + e.lane) * 2] = gm;
    L.st[(size_t)(e.wrow0 + e.lane) * 2 + 1] = rstd;
  }
}

// y = (v - mean) * rstd * gamma + beta (+ elu) on a sub-tile whose columns start at cb
__device__ __forceinline__ void ln_apply(const EP& L, int cb, float gm, float rstd, int act, float (&v)[32]) {
  const float4* gp = reinterpret_cast<const float4*>(L.gamma + cb);   // 128-byte aligned slices
  const float4* bp = reinterpret_cast<const float4*>(L.beta + cb);
#pragma unroll
  for (int q = 0; q < 8; ++q) {
    const float4 g4 = __ldg(gp + q), b4 = __ldg(bp + q);
    const float y0 = (v[4 * q] - gm) * rstd * g4.x + b4.x;
    const float y1 = (v[4 * q + 1] - gm) * rstd * g4.y + b4.y;
    const float y2 = (v[4 * q + 2] - gm) * rstd * g4.z + b4.z;
    const float y3 = (v[4 * q + 3] - gm) * rstd * g4.w + b4.w;
    v[4 * q] = act ? elu_f(y0) : y0;
    v[4 * q + 1] = act ? elu_f(y1) : y1;
    v[4 * q + 2] = act ? elu_f(y2) : y2;
    v[4 * q + 3] = act ? elu_f(y3) : y3;
  }
}

// E_PLAIN: out = acc (+ add) (+ bias)
template <int KS>
__device__ void epi_plain(const EP& L, Epi& e) {
  const bool own = owns<KS>(L, e);
  for (int c = 0; c < L.nchunk; ++c) {
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      if (h < nsub<KS>(L) && mine<KS>(e, h)) {
        float v[32];
        sub_fetch<KS>(L, e, h, v);
        if (!own) continue;
        const int cb = sub_col<KS>(L, e, c, h);
        const int cols_valid = min(32, L.N - cb);
        if (cols_valid > 0) {
          add_inputs(L, e, cb, cols_valid, v);
          store_with_lo(L, e, v, cb, cols_valid);
        }
      }
    }
  }
}

// E_LN: z = acc (+add)(+bias); out = act(LayerNorm(z) * gamma + beta)   (A/nets.py:585-602)
template <int KS>
__device__ void epi_ln_fwd(const EP& L, Epi& e) {
  const int ns = nsub<KS>(L);
  const bool own = owns<KS>(L, e);            // a pure split-K contributor only feeds its partner
  const bool single = L.nchunk == 1;          // values stay in registers between the two passes
  float keep[2][32];
  float mean = 0.f, m2 = 0.f, cnt = 0.f;
  for (int c = 0; c < L.nchunk; ++c) {
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      if (h < ns && mine<KS>(e, h)) {
        float v[32];
        sub_fetch<KS>(L, e, h, v);
        if (!own) continue;
        const int cb = sub_col<KS>(L, e, c, h);
        add_inputs(L, e, cb, 32, v);
        if (e.quad == 2 && h == 0 && c == 0) trace_ev(e.cx, 2, 16 * e.cx.pair + 2);
        if (L.z != nullptr && !single)
          sub_store(e.stg, v, L.z + (size_t)e.wrow0 * L.ld_z + cb, L.ld_z, e.rows_valid, 32, e.lane);
        if (e.quad == 2 && h == 0 && c == 0) trace_ev(e.cx, 2, 16 * e.cx.pair + 3);
        welford32(v, mean, m2, cnt);
        if (single) {
#pragma unroll
          for (int j = 0; j < 32; ++j) keep[h][j] = v[j];
        }
      }
    }
  }
  if (!own) return;
  float gm, rstd;
  if (e.quad == 2) trace_ev(e.cx, 2, 16 * e.cx.pair + 4);
  row_send<KS>(L, e, mean, m2, cnt, true);
  if (single && L.z != nullptr) {        // the pre-norm values go out while the partials travel
#pragma unroll
    for (int h = 0; h < 2; ++h)
      if (h < ns && mine<KS>(e, h))
        sub_store(e.stg, keep[h], L.z + (size_t)e.wrow0 * L.ld_z + sub_col<KS>(L, e, 0, h), L.ld_z,
                  e.rows_valid, 32, e.lane);
  }
  ln_recv<KS>(L, e, cnt, gm, rstd);
  if (e.quad == 2) trace_ev(e.cx, 2, 16 * e.cx.pair + 5);
  for (int c = 0; c < L.nchunk; ++c) {
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      if (h < ns && mine<KS>(e, h)) {
        const int cb = sub_col<KS>(L, e, c, h);
        float v[32];
        if (single) {
#pragma unroll
          for (int j = 0; j < 32; ++j) v[j] = keep[h][j];
        } else {
          sub_load(e.stg, v, L.z + (size_t)e.wrow0 * L.ld_z + cb, L.ld_z, e.rows_valid, 32, e.lane);
        }
        ln_apply(L, cb, gm, rstd, L.act, v);
        if (e.quad == 2 && h == 0 && c == 0) trace_ev(e.cx, 2, 16 * e.cx.pair + 6);
        store_with_lo(L, e, v, cb, 32);
        if (e.quad == 2 && h == 0 && c == 0) trace_ev(e.cx, 2, 16 * e.cx.pair + 7);
      }
    }
  }
}

// E_LNBWD: dh = acc (+add); dz = LayerNorm'(act'(.) dh) given the saved pre-norm z and (mean, rstd)
// dgamma += sum_rows dy * xhat, dbeta += sum_rows dy  (optional)
template <int KS>
__device__ void epi_ln_bwd(const EP& L, Epi& e) {
  const int ns = nsub<KS>(L);
  const bool own = owns<KS>(L, e);
  const bool single = L.nchunk == 1;
  const bool rowok = e.lane < e.rows_valid;
  float mean = 0.f, rstd = 0.f;
  if (rowok && own) {
    mean = L.st[(size_t)(e.wrow0 + e.lane) * 2];
    rstd = L.st[(size_t)(e.wrow0 + e.lane) * 2 + 1];
  }
  float keep[2][32];
  float s1 = 0.f, s2 = 0.f;
  for (int c = 0; c < L.nchunk; ++c) {
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      if (h < ns && mine<KS>(e, h)) {
        float v[32];
        sub_fetch<KS>(L, e, h, v);
        if (!own) continue;
        const int cb = sub_col<KS>(L, e, c, h);
        add_inputs(L, e, cb, 32, v);
        float z[32];
        sub_load(e.stg, z, L.z + (size_t)e.wrow0 * L.ld_z + cb, L.ld_z, e.rows_valid, 32, e.lane);
        float dyv[32];
        const float* gp = L.gamma + cb;
        const float* bp = L.beta + cb;
        const int act = L.act;
#pragma unroll
        for (int j = 0; j < 32; ++j) {
          const float g = __ldg(gp + j);
          const float xh = (z[j] - mean) * rstd;
          float dy = rowok ? v[j] : 0.f;
          if (act) {
            const float y = xh * g + __ldg(bp + j);
            dy = y > 0.f ? dy : dy * __expf(y);
          }
          const float dx = dy * g;
          s1 += dx;
          s2 += dx * xh;
          v[j] = dx;
          z[j] = dy * xh;
          dyv[j] = dy;
        }
        if (L.dgamma != nullptr) {       // column sums over the warp's rows, one atomic per column
          warp_colsum32(z, e.lane);
          warp_colsum32(dyv, e.lane);
          atomicAdd(L.dgamma + cb + e.lane, z[0]);
          atomicAdd(L.dbeta + cb + e.lane, dyv[0]);
        }
        if (single) {
#pragma unroll
          for (int j = 0; j < 32; ++j) keep[h][j] = v[j];
        } else {
          sub_store(e.stg, v, L.out + (size_t)e.wrow0 * L.ld_out + cb, L.ld_out, e.rows_valid, 32, e.lane);
        }
      }
    }
  }
  if (!own) return;
  float2 parts[Geo<KS>::SLOTS];
  float unused_cnt = 0.f;
  row_exchange<KS>(L, e, s1, s2, unused_cnt, false, parts);
  const int nsl = n_slots<KS>(L);
  float t1 = 0.f, t2 = 0.f;
  for (int d = 0; d < nsl; ++d) { t1 += parts[d].x; t2 += parts[d].y; }
  t1 /= (float)L.N;
  t2 /= (float)L.N;
  for (int c = 0; c < L.nchunk; ++c) {
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      if (h < ns && mine<KS>(e, h)) {
        const int cb = sub_col<KS>(L, e, c, h);
        float v[32], z[32];
        if (single) {
#pragma unroll
          for (int j = 0; j < 32; ++j) v[j] = keep[h][j];
        } else {
          sub_load(e.stg, v, L.out + (size_t)e.wrow0 * L.ld_out + cb, L.ld_out, e.rows_valid, 32, e.lane);
        }
        sub_load(e.stg, z, L.z + (size_t)e.wrow0 * L.ld_z + cb, L.ld_z, e.rows_valid, 32, e.lane);
#pragma unroll
        for (int j = 0; j < 32; ++j) {
          const float xh = (z[j] - mean) * rstd;
          v[j] = rstd * (v[j] - t1 - xh * t2);
        }
        store_with_lo(L, e, v, cb, 32);
      }
    }
  }
}

// E_GRU: joint LayerNorm over the 3 D gate pre-activations, then the GRU update (A/nets.py:149-160)
//   y = LN(z) * gamma + beta; r = sig(y_r); c = tanh(r * y_c); u = sig(y_u - 1); out = u c + (1 - u) prev
// Chunk g of a CTA holds its column slice of gate g (chunk_stride = D), so a thread has the three
// gate values of its columns.  z: pre-norm (saved), x1: y (saved), x0: prev.
template <int KS>
__device__ void epi_gru(const EP& L, Epi& e) {
  const int ns = nsub<KS>(L);
  const bool own = owns<KS>(L, e);
  const int D = L.chunk_stride;
  float mean = 0.f, m2 = 0.f, cnt = 0.f;
  for (int c = 0; c < 3; ++c) {
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      if (h < ns && mine<KS>(e, h)) {
        float v[32];
        sub_fetch<KS>(L, e, h, v);
        if (!own) continue;
        const int cb = sub_col<KS>(L, e, c, h);
        add_inputs(L, e, cb, 32, v);
        sub_store(e.stg, v, L.z + (size_t)e.wrow0 * L.ld_z + cb, L.ld_z, e.rows_valid, 32, e.lane);
        welford32(v, mean, m2, cnt);
      }
    }
  }
  if (!own) return;
  float gm, rstd;
  if (e.warp == 2) trace_ev(e.cx, 2, 16 * e.cx.pair + 12);
  ln_finish<KS>(L, e, mean, m2, cnt, gm, rstd);
  if (e.warp == 2) trace_ev(e.cx, 2, 16 * e.cx.pair + 13);
  for (int h = 0; h < ns; ++h) {
    if (!mine<KS>(e, h)) continue;
    const int cj = sub_col<KS>(L, e, 0, h);        // column inside each gate
    float g[32], v[32];
    // reset gate
    sub_load(e.stg, v, L.z + (size_t)e.wrow0 * L.ld_z + cj, L.ld_z, e.rows_valid, 32, e.lane);
    ln_apply(L, cj, gm, rstd, 0, v);
    if (L.x[1] != nullptr)
      sub_store(e.stg, v, L.x[1] + (size_t)e.wrow0 * L.ldx[1] + cj, L.ldx[1], e.rows_valid, 32, e.lane);
#pragma unroll
    for (int j = 0; j < 32; ++j) g[j] = sigmoid_f(v[j]);
    if (e.warp == 2) trace_ev(e.cx, 2, 16 * e.cx.pair + 14);
    // candidate
    sub_load(e.stg, v, L.z + (size_t)e.wrow0 * L.ld_z + D + cj, L.ld_z, e.rows_valid, 32, e.lane);
    ln_apply(L, D + cj, gm, rstd, 0, v);
    if (L.x[1] != nullptr)
      sub_store(e.stg, v, L.x[1] + (size_t)e.wrow0 * L.ldx[1] + D + cj, L.ldx[1], e.rows_valid, 32, e.lane);
#pragma unroll
    for (int j = 0; j < 32; ++j) g[j] = tanh_f(g[j] * v[j]);
    if (e.warp == 2) trace_ev(e.cx, 2, 16 * e.cx.pair + 15);
    // update gate and the new state
    sub_load(e.stg, v, L.z + (size_t)e.wrow0 * L.ld_z + 2 * D + cj, L.ld_z, e.rows_valid, 32, e.lane);
    ln_apply(L, 2 * D + cj, gm, rstd, 0, v);
    if (L.x[1] != nullptr)
      sub_store(e.stg, v, L.x[1] + (size_t)e.wrow0 * L.ldx[1] + 2 * D + cj, L.ldx[1], e.rows_valid, 32, e.lane);
    float pv[32];
    sub_load(e.stg, pv, L.x[0] + (size_t)e.wrow0 * L.ldx[0] + cj, L.ldx[0], e.rows_valid, 32, e.lane);
#pragma unroll
    for (int j = 0; j < 32; ++j) {
      const float u = sigmoid_f(v[j] - 1.f);
      v[j] = u * g[j] + (1.f - u) * pv[j];
    }
    store_with_lo(L, e, v, cj, 32);
  }
}

// E_STATS: categorical head.  z = acc + bias; logit = log((1 - mix) softmax(z) + mix / C) per group of
// C classes (A/nets.py:162-171); with noise: idx = argmax(logit + G), one-hot into x1
// (OneHotDist.sample, A/tfutils.py:368-382).  C == 32: one group per 32-column sub-tile; C < 32: the
// whole (narrow) layer is one group.  out = logit.
template <int KS>
__device__ void epi_stats(const EP& L, Epi& e) {
  const bool own = owns<KS>(L, e);
  const int C = L.ipar;
  const float mix = L.f[0], uni = mix / (float)C;
  for (int c = 0; c < L.nchunk; ++c) {
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      if (h < nsub<KS>(L) && mine<KS>(e, h)) {
        float v[32];
        sub_fetch<KS>(L, e, h, v);
        if (!own) continue;
        const int cb = sub_col<KS>(L, e, c, h);
        const int cols_valid = min(32, L.N - cb);
        if (cols_valid <= 0) continue;
        add_inputs(L, e, cb, cols_valid, v);
        if (e.warp == 2 && c == 0) trace_ev(e.cx, 2, 16 * e.cx.pair + 12);
        float mx = -INFINITY;
#pragma unroll
        for (int j = 0; j < 32; ++j) if (j < C) mx = fmaxf(mx, v[j]);
        float sum = 0.f;
#pragma unroll
        for (int j = 0; j < 32; ++j) { v[j] = j < C ? expf(v[j] - mx) : 0.f; sum += v[j]; }
        const float inv = 1.f / sum;
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] = logf((1.f - mix) * v[j] * inv + uni);
        if (e.warp == 2 && c == 0) trace_ev(e.cx, 2, 16 * e.cx.pair + 13);
        sub_store(e.stg, v, L.out + (size_t)e.wrow0 * L.ld_out + cb, L.ld_out, e.rows_valid, cols_valid, e.lane);
        if (e.warp == 2 && c == 0) trace_ev(e.cx, 2, 16 * e.cx.pair + 14);
        if (L.x[1] != nullptr) {
          if (L.x[0] != nullptr) {
            float g[32];
            sub_load(e.stg, g, L.x[0] + (size_t)e.wrow0 * L.ldx[0] + cb, L.ldx[0], e.rows_valid, cols_valid, e.lane);
#pragma unroll
            for (int j = 0; j < 32; ++j) v[j] += g[j];
          }
          float bv = -INFINITY;
          int bi = 0;
#pragma unroll
          for (int j = 0; j < 32; ++j) if (j < C && v[j] > bv) { bv = v[j]; bi = j; }   // first maximum
#pragma unroll
          for (int j = 0; j < 32; ++j) v[j] = (j == bi) ? 1.f : 0.f;
          sub_store(e.stg, v, L.x[1] + (size_t)e.wrow0 * L.ldx[1] + cb, L.ldx[1], e.rows_valid, cols_valid, e.lane);
          if (e.warp == 2 && c == 0) trace_ev(e.cx, 2, 16 * e.cx.pair + 15);
        }
      }
    }
  }
}

// one row of the action head with compile-time width (register indices stay static)
template <int A>
__device__ __forceinline__ void actor_row_vec(const float (&v)[32], float lo, float hi, const float* pe,
                                              float* pm, float* ps, float* po, float* pl) {
#pragma unroll
  for (int q = 0; q < A / 4; ++q) {
    const float4 ep = *reinterpret_cast<const float4*>(pe + 4 * q);
    const float epv[4] = {ep.x, ep.y, ep.z, ep.w};
    float mu[4], sd[4], a[4], al[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      mu[i] = tanh_f(v[4 * q + i]);
      sd[i] = (hi - lo) * sigmoid_f(v[A + 4 * q + i]) + lo;
      a[i] = mu[i] + sd[i] * epv[i];
      al[i] = lo_of(a[i]);
    }
    *reinterpret_cast<float4*>(pm + 4 * q) = make_float4(mu[0], mu[1], mu[2], mu[3]);
    *reinterpret_cast<float4*>(ps + 4 * q) = make_float4(sd[0], sd[1], sd[2], sd[3]);
    *reinterpret_cast<float4*>(po + 4 * q) = make_float4(a[0], a[1], a[2], a[3]);
    if (pl != nullptr) *reinterpret_cast<float4*>(pl + 4 * q) = make_float4(al[0], al[1], al[2], al[3]);
  }
}

// E_ACTOR: continuous action head (A/nets.py:461-468), N = 2 A columns [mean raw | std raw]:
// mean = tanh(m); std = (hi - lo) sigmoid(s) + lo; action = mean + std * eps.
// out = action (+ out_lo), x0 = eps, x1 = mean, x2 = std.
template <int KS>
__device__ void epi_actor(const EP& L, Epi& e) {
  const int A = L.ipar;
  float v[32];
  sub_fetch<KS>(L, e, 0, v);
  if (!owns<KS>(L, e)) return;
  add_inputs(L, e, 0, 2 * A, v);
  if (e.lane < e.rows_valid) {
    const size_t row = (size_t)(e.wrow0 + e.lane);
    const float lo = L.f[0], hi = L.f[1];
    float* po = L.out + row * L.ld_out;
    float* pl = L.out_lo != nullptr ? L.out_lo + row * L.ld_lo : nullptr;
    float* pm = L.x[1] + row * L.ldx[1];
    float* ps = L.x[2] + row * L.ldx[2];
    const float* pe = L.x[0] + row * L.ldx[0];
    const bool vec = (A & 3) == 0 && ((L.ld_out | L.ldx[0] | L.ldx[1] | L.ldx[2] | (pl ? L.ld_lo : 0)) & 3) == 0 &&
                     ((reinterpret_cast<uintptr_t>(L.out) | reinterpret_cast<uintptr_t>(L.x[0]) |
                       reinterpret_cast<uintptr_t>(L.x[1]) | reinterpret_cast<uintptr_t>(L.x[2]) |
                       reinterpret_cast<uintptr_t>(L.out_lo)) & 15) == 0;
    if (vec) {                       // 16-byte accesses: a row of the head is 48 bytes at A = 12
      if (A == 4) actor_row_vec<4>(v, lo, hi, pe, pm, ps, po, pl);
      else if (A == 8) actor_row_vec<8>(v, lo, hi, pe, pm, ps, po, pl);
      else if (A == 12) actor_row_vec<12>(v, lo, hi, pe, pm, ps, po, pl);
      else actor_row_vec<16>(v, lo, hi, pe, pm, ps, po, pl);
    } else {
#pragma unroll
      for (int j = 0; j < 16; ++j) {
        if (j < A) {
          const float mu = tanh_f(v[j]);
          float sraw = 0.f;
#pragma unroll
          for (int t = 0; t < 32; ++t) if (t == A + j) sraw = v[t];
          const float sd = (hi - lo) * sigmoid_f(sraw) + lo;
          const float a = mu + sd * pe[j];
          pm[j] = mu;
          ps[j] = sd;
          po[j] = a;
          if (pl != nullptr) pl[j] = lo_of(a);
        }
      }
    }
  }
}


// E_STATSBWD: gradient of the straight-through categorical sample back to the raw logits
// (reverse of E_STATS with dLogit = 0; A/tfutils.py:375-381, A/nets.py:166-170).  input (acc + add) =
// dStoch, x0 = the unimixed logits saved by the forward pass, f0 = mix, ipar = classes (32).
//   q = exp(lg); dl = q (ds - sum q ds); dq = dl / q; p = (q - mix/C) / (1 - mix);
//   dz = p (dp - sum p dp),  dp = (1 - mix) dq
template <int KS>
__device__ void epi_stats_bwd(const EP& L, Epi& e) {
  const bool own = owns<KS>(L, e);
  const int C = L.ipar;
  const float mix = L.f[0], uni = mix / (float)C;
  for (int c = 0; c < L.nchunk; ++c) {
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      if (h < nsub<KS>(L) && mine<KS>(e, h)) {
        float v[32];
        sub_fetch<KS>(L, e, h, v);
        if (!own) continue;
        const int cb = sub_col<KS>(L, e, c, h);
        add_inputs(L, e, cb, 32, v);
        float lg[32];
        sub_load(e.stg, lg, L.x[0] + (size_t)e.wrow0 * L.ldx[0] + cb, L.ldx[0], e.rows_valid, 32, e.lane);
        float sq = 0.f;
#pragma unroll
        for (int j = 0; j < 32; ++j) { lg[j] = expf(lg[j]); sq += lg[j] * v[j]; }     // lg := q
        float sp = 0.f;
#pragma unroll
        for (int j = 0; j < 32; ++j) {
          const float q = lg[j];
          const float dq = v[j] - sq;                      // dl / q
          const float pj = (q - uni) / (1.f - mix);
          const float dp = (1.f - mix) * dq;
          lg[j] = pj;
          v[j] = dp;
          sp += pj * dp;
        }
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] = lg[j] * (v[j] - sp);
        store_with_lo(L, e, v, cb, 32);
      }
    }
  }
}

// E_GRUBWD: reverse of E_GRU.  Layer width N = D: input (acc + add) = dOut, the gradient wrt the new
// deter.  x0 = previous deter, x1 = y (post-norm gates saved by the forward pass), z / st = the
// pre-norm gates and their (mean, rstd), gamma / beta over 3 D.  Outputs: out [M, 3 D] = dz of the
// joint LayerNorm (+ out_lo), x2 [M, D] += dOut (1 - u)  (gradient wrt the previous deter).
template <int KS>
__device__ void epi_gru_bwd(const EP& L, Epi& e) {
  const int ns = nsub<KS>(L);
  const bool own = owns<KS>(L, e);
  const int D = L.N;
  const bool rowok = e.lane < e.rows_valid;
  float mean = 0.f, rstd = 0.f;
  if (rowok && own) {
    mean = L.st[(size_t)(e.wrow0 + e.lane) * 2];
    rstd = L.st[(size_t)(e.wrow0 + e.lane) * 2 + 1];
  }
  float s1 = 0.f, s2 = 0.f;
  for (int c = 0; c < L.nchunk; ++c) {
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      if (h < ns && mine<KS>(e, h)) {
        float v[32];
        sub_fetch<KS>(L, e, h, v);
        if (!own) continue;
        const int cj = sub_col<KS>(L, e, c, h);          // column inside D
        add_inputs(L, e, cj, 32, v);
        float yr[32], yc[32], t[32];
        sub_load(e.stg, yr, L.x[1] + (size_t)e.wrow0 * L.ldx[1] + cj, L.ldx[1], e.rows_valid, 32, e.lane);
        sub_load(e.stg, yc, L.x[1] + (size_t)e.wrow0 * L.ldx[1] + D + cj, L.ldx[1], e.rows_valid, 32, e.lane);
        sub_load(e.stg, t, L.x[1] + (size_t)e.wrow0 * L.ldx[1] + 2 * D + cj, L.ldx[1], e.rows_valid, 32, e.lane);
        float pv[32];
        sub_load(e.stg, pv, L.x[0] + (size_t)e.wrow0 * L.ldx[0] + cj, L.ldx[0], e.rows_valid, 32, e.lane);
        // gates: yr := d y_r, yc := d y_c, t := d y_u, pv := d prev
#pragma unroll
        for (int j = 0; j < 32; ++j) {
          const float dout = rowok ? v[j] : 0.f;
          const float r = sigmoid_f(yr[j]);
          const float cc = tanh_f(r * yc[j]);
          const float u = sigmoid_f(t[j] - 1.f);
          const float du = dout * (cc - pv[j]);
          const float dpre = dout * u * (1.f - cc * cc);
          yr[j] = dpre * yc[j] * r * (1.f - r);
          yc[j] = dpre * r;
          t[j] = du * u * (1.f - u);
          pv[j] = dout * (1.f - u);
        }
        if (L.x[2] != nullptr) {                     // d prev accumulates into the caller's buffer
          float old[32];
          sub_load(e.stg, old, L.x[2] + (size_t)e.wrow0 * L.ldx[2] + cj, L.ldx[2], e.rows_valid, 32, e.lane);
#pragma unroll
          for (int j = 0; j < 32; ++j) pv[j] += old[j];
          sub_store(e.stg, pv, L.x[2] + (size_t)e.wrow0 * L.ldx[2] + cj, L.ldx[2], e.rows_valid, 32, e.lane);
        }
        // joint LayerNorm backward, pass 1: dx = dy * gamma, row sums; dx parked in `out`
#pragma unroll
        for (int g = 0; g < 3; ++g) {
          float (&dy)[32] = g == 0 ? yr : (g == 1 ? yc : t);
          float z[32];
          sub_load(e.stg, z, L.z + (size_t)e.wrow0 * L.ld_z + g * D + cj, L.ld_z, e.rows_valid, 32, e.lane);
          const float* gp = L.gamma + g * D + cj;
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            const float dx = dy[j] * __ldg(gp + j);
            const float xh = (z[j] - mean) * rstd;
            s1 += dx;
            s2 += dx * xh;
            dy[j] = dx;
          }
          sub_store(e.stg, dy, L.out + (size_t)e.wrow0 * L.ld_out + g * D + cj, L.ld_out, e.rows_valid, 32, e.lane);
        }
      }
    }
  }
  if (!own) return;
  float2 parts[Geo<KS>::SLOTS];
  float unused_cnt = 0.f;
  row_exchange<KS>(L, e, s1, s2, unused_cnt, false, parts);
  const int nsl = n_slots<KS>(L);
  float t1 = 0.f, t2 = 0.f;
  for (int d = 0; d < nsl; ++d) { t1 += parts[d].x; t2 += parts[d].y; }
  t1 /= (float)(3 * D);
  t2 /= (float)(3 * D);
  for (int c = 0; c < L.nchunk; ++c) {
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      if (h < ns && mine<KS>(e, h)) {
        const int cj = sub_col<KS>(L, e, c, h);
#pragma unroll
        for (int g = 0; g < 3; ++g) {
          float v[32], z[32];
          sub_load(e.stg, v, L.out + (size_t)e.wrow0 * L.ld_out + g * D + cj, L.ld_out, e.rows_valid, 32, e.lane);
          sub_load(e.stg, z, L.z + (size_t)e.wrow0 * L.ld_z + g * D + cj, L.ld_z, e.rows_valid, 32, e.lane);
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            const float xh = (z[j] - mean) * rstd;
            v[j] = rstd * (v[j] - t1 - xh * t2);
          }
          store_with_lo(L, e, v, g * D + cj, 32);
        }
      }
    }
  }
}

template <int KS>
__device__ void epilogue(const Shared& sh, const LmLayer* layers, int nlayers, int M, int rank,
                         int cluster_id, int nclusters, int warp, int lane, uint32_t tmem_base,
                         Ctx cx) {
  constexpr int CLS = Geo<KS>::CLS, EW = Geo<KS>::EW;
  const int ntiles = (M + BM - 1) / BM;
  Epi e;
  e.sh = sh; e.cx = cx; e.tmem_base = tmem_base;
  e.ac = 0; e.xc = 0; e.rc = 0;
  e.rank = rank; e.cs = rank % NS; e.kh = rank / NS;
  e.warp = warp; e.quad = warp & 3; e.lane = lane;      // quad: the TMEM lane quadrant this warp may read
  e.ew = warp < 6 ? warp - 2 : warp - 3;                // warps 2-5 -> 0-3, warps 7-10 -> 4-7
  e.grp = e.ew >> 2;
  e.stg = reinterpret_cast<float*>(sh.smem + Geo<KS>::OFF_STG + e.ew * STG_WARP_BYTES);
  e.stats_off = Geo<KS>::OFF_STATS;
  e.red_off = Geo<KS>::OFF_RED;
  e.loc_off = Geo<KS>::OFF_LOC;
  EP* ep_slot = reinterpret_cast<EP*>(sh.smem + Geo<KS>::OFF_EP + e.ew * EP_BYTES);
  int pair = 0;
  int base0 = 0, base1 = 0;                 // layers the two chains completed in earlier tile groups
  const bool dual = dual_mode(ntiles, nclusters, cx.flags);
  const int tstep = dual ? 2 * nclusters : nclusters;
  e.trow = e.quad * 32 + lane;
  for (int t0 = cluster_id; t0 < ntiles; t0 += tstep) {
    const int ng = (dual && t0 + nclusters < ntiles) ? 2 : 1;
    for (int l = 0; l < nlayers; ++l) {
     for (int sub = 0; sub < ng; ++sub, ++pair) {
      e.wrow0 = (t0 + sub * nclusters) * BM + e.quad * 32;      // first global row of this warp
      e.rows_valid = max(0, min(32, M - e.wrow0));
      const uint32_t ready = sh.bars + BAR_READY + 4 * sub;     // this chain's counter
      __syncwarp();
      if (lane == 0) *ep_slot = load_ep(layers[l]);
      __syncwarp();
      const EP& L = *ep_slot;
      e.cx.pair = pair;
      const bool active = e.cs < L.nactive && (KS == 1 || L.bn == 64 || e.kh == 0);
      if (L.nseg == 0 && l > 0)       // no producer wait protects its inputs: wait here
        count_wait(ready, (uint32_t)((sub ? base1 : base0) + l) * CLS, e.cx, 8);
      if (warp == 2) trace_ev(e.cx, 2, 16 * pair);
      if (active && e.grp >= nsub<KS>(L)) {
        // a 32-wide slice has no second half: this warp group only hands the accumulator back
        if (has_k(L, e, KS)) {
          for (int c = 0; c < L.nchunk; ++c) { acc_wait(e); acc_release(e); }
        }
      } else if (active) {
        if (L.epi == DD_EPI_PLAIN) epi_plain<KS>(L, e);
        else if (L.epi == DD_EPI_LN) epi_ln_fwd<KS>(L, e);
        else if (L.epi == DD_EPI_LNBWD) epi_ln_bwd<KS>(L, e);
        else if (L.epi == DD_EPI_GRU) epi_gru<KS>(L, e);
        else if (L.epi == DD_EPI_STATS) epi_stats<KS>(L, e);
        else if (L.epi == DD_EPI_ACTOR) epi_actor<KS>(L, e);
        else if (L.epi == DD_EPI_STATSBWD) epi_stats_bwd<KS>(L, e);
        else epi_gru_bwd<KS>(L, e);
      }
      // this CTA's rows of the layer's outputs are written: tell every CTA's producer.  Each
      // writer orders its stores against the async proxy (TMA reads them), the epilogue warps
      // meet at a named barrier, then one thread publishes with a single cluster-scope fence.
      if (warp == 2) trace_ev(e.cx, 2, 16 * pair + 9);
      asm volatile("fence.proxy.async;" ::: "memory");
      asm volatile("bar.sync 1, %0;" ::"n"(EW * 32) : "memory");
      if (warp == 2 && lane == 0) {
        asm volatile("fence.acq_rel.cluster;" ::: "memory");
#pragma unroll
        for (int d = 0; d < CLS; ++d) count_add_remote(ready, d);
      }
      if (warp == 2) trace_ev(e.cx, 2, 16 * pair + 11);
     }
    }
    base0 += nlayers;
    if (ng == 2) base1 += nlayers;
  }
}

template <int KS>
__global__ void __launch_bounds__(Geo<KS>::THREADS, 1)
lm_kernel(const LmLayer* __restrict__ layers, int nlayers, int M, unsigned* dbg,
          unsigned long long* prof, int flags) {
  using G = Geo<KS>;
  extern __shared__ __align__(1024) char smem[];
  // warp index and cluster rank broadcast from lane 0: the role dispatch and everything steered by
  // them compiles as warp-uniform control flow
  const int tid = threadIdx.x, warp = __shfl_sync(0xffffffffu, tid >> 5, 0), lane = tid & 31;
  const int rank = __shfl_sync(0xffffffffu, (int)cluster_rank(), 0);
  const int cluster_id = blockIdx.x / G::CLS, nclusters = gridDim.x / G::CLS;
  Shared sh;
  sh.smem = smem;
  sh.base = smem_u32(smem);
  sh.bars = sh.base + G::OFF_BAR;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + G::OFF_BAR + BAR_TMEM);

  if (tid == 0) {
    for (int s = 0; s < G::STAGES; ++s) {
      mbar_init(sh.bars + BAR_FULL + 8 * s, 2);       // one arrive.expect_tx per producer warp
      mbar_init(sh.bars + BAR_EMPTY + 8 * s, NS);     // the CTAs that receive my activation slices
    }
    for (int b = 0; b < 2; ++b) {
      mbar_init(sh.bars + BAR_ACCF + 8 * b, 1);
      mbar_init(sh.bars + BAR_ACCE + 8 * b, G::EW);     // one arrive per epilogue warp
    }
    for (int b = 0; b < 2; ++b) {
      mbar_init(sh.bars + BAR_STATS + 8 * b, 1);
      mbar_init(sh.bars + BAR_RED + 8 * b, 1);
    }
    *reinterpret_cast<volatile unsigned long long*>(smem + G::OFF_BAR + BAR_READY) = 0ull;
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;"
                 ::"r"(smem_u32(tmem_slot)), "r"((uint32_t)TMEM_COLS) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  cluster_sync_all();                // every CTA's barriers exist before anyone multicasts
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = *tmem_slot;
  const long long kstart = clock64();
  Ctx cx = {dbg, 0, prof != nullptr ? prof + (size_t)blockIdx.x * 16 : nullptr, flags,
            (prof != nullptr && blockIdx.x == 0) ? prof + 148 * 16 : nullptr};
  if (tid == 0) trace_ev(cx, 3, 0);
  if ((sh.base & 1023u) != 0u) stuck(cx, 9, sh.base);      // SWIZZLE_128B tiles need 1 KB alignment

  if (warp == 0) {
    producer_b<KS>(sh, layers, nlayers, M, rank, cluster_id, nclusters, lane, cx);
  } else if (warp == 6) {
    producer_a<KS>(sh, layers, nlayers, M, rank, cluster_id, nclusters, lane, cx);
  } else if (warp == 1) {
    mma_issuer<KS>(sh, layers, nlayers, M, rank, cluster_id, nclusters, lane, tmem_base, cx);
  } else {
    epilogue<KS>(sh, layers, nlayers, M, rank, cluster_id, nclusters, warp, lane, tmem_base, cx);
  }

  if (tid % 32 == 0 && warp < 6) trace_ev(cx, 3, 1 + warp);
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (tid == 0) trace_ev(cx, 3, 8);
  cluster_sync_all();                // nobody exits while a peer may still write into it
  if (tid == 0) trace_ev(cx, 3, 9);
  if (tid == 0 && cx.prof != nullptr) cx.prof[0] += (unsigned long long)(clock64() - kstart);
  if (warp == 1) {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;"
                 ::"r"(tmem_base), "r"((uint32_t)TMEM_COLS) : "memory");
  }
}

// ---- operand preparation ------------------------------------------------------------------
__global__ void split_lo_kernel(const float* __restrict__ x, int ldx, float* __restrict__ lo,
                                int ldlo, int rows, int cols) {
  const int64_t n = (int64_t)rows * cols;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n;
       i += (int64_t)gridDim.x * blockDim.x) {
    const int r = (int)(i / cols), c = (int)(i % cols);
    lo[(size_t)r * ldlo + c] = lo_of(x[(size_t)r * ldx + c]);
  }
}

struct TrJobs {
  int n;
  int src[DD_MAX_TRANSPOSE_JOBS], rows[DD_MAX_TRANSPOSE_JOBS], cols[DD_MAX_TRANSPOSE_JOBS];
  int dst[DD_MAX_TRANSPOSE_JOBS], ldd[DD_MAX_TRANSPOSE_JOBS], mode[DD_MAX_TRANSPOSE_JOBS];
  int tile0[DD_MAX_TRANSPOSE_JOBS + 1];
};

// per job, for the row-major [rows, cols] matrix at p + src:
//   mode 0 (transpose): pt[dst + c*ldd + r] = p[src + r*cols + c]
//   mode 1 (copy):      pt[dst + r*ldd + c] = p[src + r*cols + c]     (re-strided copy)
// and the tf32 remainder of the same element into pt_lo; 32 x 32 tiles
__global__ void transpose_lo_kernel(const float* __restrict__ p, float* __restrict__ pt,
                                    float* __restrict__ pt_lo, TrJobs jobs) {
  __shared__ float t[32][33];
  int j = 0;
  while (j + 1 < jobs.n && (int)blockIdx.x >= jobs.tile0[j + 1]) ++j;
  const int rows = jobs.rows[j], cols = jobs.cols[j], ldd = jobs.ldd[j];
  const int tiles_c = (cols + 31) / 32;
  const int tl = (int)blockIdx.x - jobs.tile0[j];
  const int r0 = (tl / tiles_c) * 32, c0 = (tl % tiles_c) * 32;
  const float* src = p + jobs.src[j];
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;       // 32 x 8
  if (jobs.mode[j] == 1) {
    for (int i = ty; i < 32; i += 8) {
      const int r = r0 + i, c = c0 + tx;
      if (r < rows && c < cols) {
        const float x = src[(size_t)r * cols + c];
        pt[(size_t)jobs.dst[j] + (size_t)r * ldd + c] = x;
        pt_lo[(size_t)jobs.dst[j] + (size_t)r * ldd + c] = lo_of(x);
      }
    }
    return;
  }
  for (int i = ty; i < 32; i += 8) {
    const int r = r0 + i, c = c0 + tx;
    t[i][tx] = (r < rows && c < cols) ? src[(size_t)r * cols + c] : 0.f;
  }
  __syncthreads();
  for (int i = ty; i < 32; i += 8) {
    const int c = c0 + i, r = r0 + tx;
    if (r < rows && c < cols) {
      const float x = t[tx][i];
      pt[(size_t)jobs.dst[j] + (size_t)c * ldd + r] = x;
      pt_lo[(size_t)jobs.dst[j] + (size_t)c * ldd + r] = lo_of(x);
    }
  }
}

// ---- host side ------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*,
                                  const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                  const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn encode_fn() {
  static EncodeTiledFn fn = nullptr;
  if (fn == nullptr) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qr;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qr) == cudaSuccess &&
        qr == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  }
  return fn;
}

// 3-D view {cols, rows, 2 planes} of a (w, w_lo) pair whose planes are `plane` bytes apart: one box
// {32, box_rows, 2} lands w rows then w_lo rows contiguously in shared memory
static int make_map3(CUtensorMap* m, const float* base, int64_t rows, int64_t cols, int64_t ld,
                     int64_t plane_bytes, int box_rows) {
  EncodeTiledFn fn = encode_fn();
  if (fn == nullptr) return fail("dd_program_build", "cuTensorMapEncodeTiled unavailable");
  cuuint64_t dims[3] = {(cuuint64_t)cols, (cuuint64_t)rows, 2};
  cuuint64_t strides[2] = {(cuuint64_t)ld * 4, (cuuint64_t)plane_bytes};
  cuuint32_t box[3] = {(cuuint32_t)BK, (cuuint32_t)box_rows, 2};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = fn(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, const_cast<float*>(base), dims, strides, box,
                  estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                  CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    char msg[64];
    snprintf(msg, sizeof(msg), "cuTensorMapEncodeTiled (3-D) failed (%d)", (int)r);
    return fail("dd_program_build", msg);
  }
  return 0;
}

// 2-D fp32 tensor [rows, cols], row stride ld floats, box {32 cols, box_rows}, SWIZZLE_128B,
// out-of-bounds elements read as zero
static int make_map(CUtensorMap* m, const float* base, int64_t rows, int64_t cols, int64_t ld,
                    int box_rows) {
  EncodeTiledFn fn = encode_fn();
  if (fn == nullptr) return fail("dd_program_build", "cuTensorMapEncodeTiled unavailable");
  if ((reinterpret_cast<uintptr_t>(base) & 15) || (ld & 3))
    return fail("dd_program_build", "operand base / row stride must be 16-byte aligned");
  cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)ld * 4};
  cuuint32_t box[2] = {(cuuint32_t)BK, (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = fn(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(base), dims, strides, box,
                  estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                  CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    char msg[64];
    snprintf(msg, sizeof(msg), "cuTensorMapEncodeTiled failed (%d)", (int)r);
    return fail("dd_program_build", msg);
  }
  return 0;
}

static unsigned* g_dbg_host = nullptr;
static unsigned* g_dbg_dev = nullptr;
static unsigned long long* g_prof_dev = nullptr;     // [148 CTAs][16] wait-cycle totals
static int g_prof_on = 0;
static int g_flags = 0;

}  // namespace lm
}  // namespace dd

using namespace dd;

extern "C" size_t dd_program_bytes(int nlayers) { return sizeof(lm::LmLayer) * (size_t)nlayers; }

namespace dd {
namespace lm {
// ---- which earlier layers write an A segment ---------------------------------------------------
// A rectangle of floats: `rows` rows of `w` floats, `ld` apart.
struct Rect { const float* p; int64_t w, ld; };
static bool rect_overlap(const Rect& a, const Rect& b, int64_t rows) {
  if (a.p == nullptr || b.p == nullptr || a.w <= 0 || b.w <= 0) return false;
  if (a.ld == b.ld && a.ld > 0) {
    // same pitch: express a's origin in b's (row, column) grid
    const int64_t d = a.p - b.p, ld = a.ld;
    int64_t q = d / ld, rem = d % ld;
    if (rem < 0) { rem += ld; q -= 1; }
    if (rem + a.w <= ld) {                               // a's rows do not wrap in b's grid
      if (q <= -rows || q >= rows) return false;          // no row in common
      return rem < b.w;                                  // column intervals [rem, rem + a.w) and [0, b.w)
    }
    return q > -rows - 1 && q < rows + 1;                // wrapping: be conservative
  }
  const float* a_end = a.p + (rows - 1) * a.ld + a.w;
  const float* b_end = b.p + (rows - 1) * b.ld + b.w;
  return a.p < b_end && b.p < a_end;
}
// everything layer s writes that a later layer could read as a GEMM operand (generous widths)
static int written_rects(const DdLayer& s, Rect* out) {
  int n = 0;
  const int64_t N = s.N;
  const int64_t ow = s.epi == DD_EPI_GRU ? N / 3 : (s.epi == DD_EPI_GRUBWD ? 3 * N : (s.epi == DD_EPI_ACTOR ? s.ipar : N));
  out[n++] = {s.out, ow, s.ld_out};
  if (s.out_lo != nullptr) out[n++] = {s.out_lo, ow, s.ld_lo};
  if (s.z != nullptr && (s.epi == DD_EPI_LN || s.epi == DD_EPI_GRU || s.epi == DD_EPI_PLAIN))
    out[n++] = {s.z, N, s.ld_z};
  if (s.st != nullptr && (s.epi == DD_EPI_LN || s.epi == DD_EPI_GRU)) out[n++] = {s.st, 2, 2};
  if (s.x[1] != nullptr && (s.epi == DD_EPI_GRU || s.epi == DD_EPI_STATS || s.epi == DD_EPI_ACTOR))
    out[n++] = {s.x[1], s.epi == DD_EPI_ACTOR ? (int64_t)s.ipar : N, s.ldx[1]};
  if (s.x[2] != nullptr && (s.epi == DD_EPI_ACTOR || s.epi == DD_EPI_GRUBWD))
    out[n++] = {s.x[2], s.epi == DD_EPI_ACTOR ? (int64_t)s.ipar : N, s.ldx[2]};
  return n;
}
// seg_dep[g] of layer l = 1 + the last earlier layer that writes any part of segment g (0: none)
static void find_segment_deps(const DdLayer* layers, int nlayers, int rows, LmLayer* host) {
  for (int l = 0; l < nlayers; ++l) {
    for (int g = 0; g < layers[l].nseg; ++g) {
      const DdSeg& sg = layers[l].seg[g];
      const Rect a = {sg.a, sg.k, sg.lda}, alo = {sg.a_lo, sg.a_lo != nullptr ? sg.k : 0, sg.lda_lo};
      int dep = 0;
      for (int j = l - 1; j >= 0 && dep == 0; --j) {
        Rect w[8];
        const int n = written_rects(layers[j], w);
        for (int i = 0; i < n; ++i)
          if (rect_overlap(a, w[i], rows) || rect_overlap(alo, w[i], rows)) { dep = j + 1; break; }
      }
      host[l].seg_dep[g] = dep;
    }
  }
}
}  // namespace lm
}  // namespace dd

// Host-only view of the dependency analysis (tests / tooling): deps[2 * l + g] = number of leading
// layers that must be complete before segment g of layer l may be fetched (0: a program input).
extern "C" int dd_program_deps(const DdLayer* layers, int nlayers, int rows, int* deps) {
  if (nlayers < 1 || rows < 1 || deps == nullptr) return fail("dd_program_deps", "empty program");
  lm::LmLayer* host = static_cast<lm::LmLayer*>(calloc(nlayers, sizeof(lm::LmLayer)));
  if (host == nullptr) return fail("dd_program_deps", "out of host memory");
  lm::find_segment_deps(layers, nlayers, rows, host);
  for (int l = 0; l < nlayers; ++l) { deps[2 * l] = host[l].seg_dep[0]; deps[2 * l + 1] = host[l].seg_dep[1]; }
  free(host);
  return 0;
}

extern "C" int dd_program_build(const DdLayer* layers, int nlayers, int rows, void* table_dev) {
  if (nlayers < 1 || rows < 1) return fail("dd_program_build", "empty program");
  lm::LmLayer* host = static_cast<lm::LmLayer*>(calloc(nlayers, sizeof(lm::LmLayer)));
  if (host == nullptr) return fail("dd_program_build", "out of host memory");
  int rc = 0;
  for (int i = 0; i < nlayers && rc == 0; ++i) {
    const DdLayer& s = layers[i];
    lm::LmLayer& d = host[i];
    if (s.nseg < 0 || s.nseg > 2) { rc = fail("dd_program_build", "0, 1 or 2 A segments"); break; }
    if (s.nseg == 0 && s.add == nullptr) { rc = fail("dd_program_build", "epilogue-only layer without input (add)"); break; }
    // column tiling across the 8 CTAs
    if (s.N % (lm::NS * 64) == 0) {
      d.bn = 64; d.nchunk = s.N / (lm::NS * 64); d.nactive = lm::NS;
    } else if (s.N % (lm::NS * 32) == 0) {
      d.bn = 32; d.nchunk = s.N / (lm::NS * 32); d.nactive = lm::NS;
    } else if (s.N <= 64 && (s.epi == DD_EPI_PLAIN || s.epi == DD_EPI_STATS || s.epi == DD_EPI_ACTOR)) {
      d.bn = s.N <= 32 ? 32 : 64; d.nchunk = 1; d.nactive = 1;
    } else {
      rc = fail("dd_program_build", "N must be a multiple of 256 (or <= 64 for plain layers)");
      break;
    }
    d.N = s.N;
    d.chunk_stride = d.nactive == 1 ? 0 : s.N / d.nchunk;
    d.nseg = s.nseg;
    for (int g = 0; g < s.nseg && rc == 0; ++g) {
      const DdSeg& sg = s.seg[g];
      if (sg.k < 1) { rc = fail("dd_program_build", "empty segment"); break; }
      d.seg_k[g] = sg.k;
      d.seg_wk0[g] = sg.wk0;
      d.seg_exact[g] = sg.a_lo == nullptr ? 1 : 0;
      rc = lm::make_map(&d.a_hi[g], sg.a, rows, sg.k, sg.lda, lm::SLICE_ROWS);
      if (rc == 0 && sg.a_lo != nullptr)
        rc = lm::make_map(&d.a_lo[g], sg.a_lo, rows, sg.k, sg.lda_lo, lm::SLICE_ROWS);
    }
    if (rc != 0) break;
    if (s.nseg > 0) {
      if (s.w == nullptr || s.w_lo == nullptr) { rc = fail("dd_program_build", "B operand needs (w, w_lo)"); break; }
      // (w, w_lo) planes a fixed, 16-byte aligned distance apart: one 3-D descriptor
      const int64_t plane = reinterpret_cast<const char*>(s.w_lo) - reinterpret_cast<const char*>(s.w);
      if (plane > 0 && (plane & 15) == 0 && plane >= (int64_t)s.wrows * s.ldw * 4 - (s.ldw - s.wcols) * 4 &&
          (reinterpret_cast<uintptr_t>(s.w) & 15) == 0 && (s.ldw & 3) == 0) {
        rc = lm::make_map3(&d.w_hi, s.w, s.wrows, s.wcols, s.ldw, plane, d.bn);
        d.w3d = 1;
      } else {
        rc = lm::make_map(&d.w_hi, s.w, s.wrows, s.wcols, s.ldw, d.bn);
        if (rc == 0) rc = lm::make_map(&d.w_lo, s.w_lo, s.wrows, s.wcols, s.ldw, d.bn);
      }
      if (rc != 0) break;
    }
    d.wn0 = s.wn0;
    d.epi = s.epi;
    d.act = s.act;
    d.ld_add = s.ld_add; d.ld_out = s.ld_out; d.ld_z = s.ld_z; d.ld_lo = s.out_lo != nullptr ? s.ld_lo : 0;
    d.add = s.add; d.bias = s.bias; d.gamma = s.gamma; d.beta = s.beta;
    d.z = s.z; d.st = s.st; d.out = s.out; d.out_lo = s.out_lo;
    d.dgamma = s.dgamma; d.dbeta = s.dbeta;
    for (int q = 0; q < 3; ++q) { d.x[q] = s.x[q]; d.ldx[q] = s.ldx[q]; }
    d.f[0] = s.f[0]; d.f[1] = s.f[1]; d.ipar = s.ipar;
    if (s.epi == DD_EPI_GRU) {
      if (d.nchunk != 3 || d.nactive != lm::NS || s.gamma == nullptr || s.beta == nullptr ||
          s.z == nullptr || s.x[0] == nullptr)
        rc = fail("dd_program_build", "GRU layer: N = 3 D with D a multiple of 256, needs scale, bias, z and prev");
    }
    if (s.epi == DD_EPI_STATS) {
      const bool wide = d.nactive == lm::NS;
      if ((wide && s.ipar != 32) || (!wide && (s.ipar != s.N || s.N > 32)) || s.nseg == 0)
        rc = fail("dd_program_build", "STATS layer: groups of 32 classes, or one narrow group of N <= 32");
    }
    if (s.epi == DD_EPI_STATSBWD) {
      if (d.nactive != lm::NS || s.ipar != 32 || s.x[0] == nullptr)
        rc = fail("dd_program_build", "STATSBWD layer: groups of 32 classes over a full-width layer, needs the logits");
    }
    if (s.epi == DD_EPI_GRUBWD) {
      if (d.nactive != lm::NS || s.gamma == nullptr || s.z == nullptr || s.st == nullptr ||
          s.x[0] == nullptr || s.x[1] == nullptr)
        rc = fail("dd_program_build", "GRUBWD layer: N = D a multiple of 256, needs scale, z, stats, prev and y");
    }
    if (s.epi == DD_EPI_ACTOR) {
      if (s.N != 2 * s.ipar || s.ipar > 16 || s.x[0] == nullptr || s.x[1] == nullptr || s.x[2] == nullptr)
        rc = fail("dd_program_build", "ACTOR layer: N = 2 A <= 32 with eps, mean and std buffers");
    }
    if (rc != 0) break;
    if (s.out == nullptr) { rc = fail("dd_program_build", "layer without output"); break; }
    if (s.epi == DD_EPI_LN && (s.gamma == nullptr || s.beta == nullptr))
      rc = fail("dd_program_build", "LayerNorm layer without scale / bias");
    if ((reinterpret_cast<uintptr_t>(s.gamma) & 15) || (reinterpret_cast<uintptr_t>(s.beta) & 15))
      rc = fail("dd_program_build", "LayerNorm scale / bias must be 16-byte aligned");
    if (s.epi == DD_EPI_LN && s.z == nullptr && !(d.nchunk == 1))
      rc = fail("dd_program_build", "multi-chunk LayerNorm needs the pre-norm buffer z");
    if (s.epi == DD_EPI_LNBWD && (s.gamma == nullptr || s.beta == nullptr || s.z == nullptr || s.st == nullptr))
      rc = fail("dd_program_build", "LayerNorm backward needs scale, bias, z and stats");
    if ((s.dgamma == nullptr) != (s.dbeta == nullptr))
      rc = fail("dd_program_build", "dgamma and dbeta come together");
  }
  if (rc == 0) lm::find_segment_deps(layers, nlayers, rows, host);
  if (rc == 0) {
    cudaError_t e = cudaMemcpy(table_dev, host, sizeof(lm::LmLayer) * (size_t)nlayers, cudaMemcpyHostToDevice);
    if (e != cudaSuccess) rc = fail("dd_program_build", cudaGetErrorString(e));
  }
  free(host);
  if (rc == 0 && lm::g_dbg_host == nullptr) {
    if (cudaHostAlloc(reinterpret_cast<void**>(&lm::g_dbg_host), 64, cudaHostAllocMapped) == cudaSuccess) {
      memset(lm::g_dbg_host, 0, 64);
      cudaHostGetDevicePointer(reinterpret_cast<void**>(&lm::g_dbg_dev), lm::g_dbg_host, 0);
    }
  }
  return rc;
}

namespace dd {
namespace lm {
static int g_ks = -1;            // 1: clusters of 8, 2: split-K clusters of 16; -1: not chosen yet
static int g_max_clusters[3] = {0, 0, 0};

template <int KS>
static int setup_variant() {
  using G = Geo<KS>;
  if (cudaFuncSetAttribute(lm_kernel<KS>, cudaFuncAttributeMaxDynamicSharedMemorySize, G::SMEM_BYTES) != cudaSuccess) {
    cudaGetLastError();
    return 0;
  }
  if (G::CLS > 8 &&
      cudaFuncSetAttribute(lm_kernel<KS>, cudaFuncAttributeNonPortableClusterSizeAllowed, 1) != cudaSuccess) {
    cudaGetLastError();
    return 0;
  }
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(G::CLS * 8);
  cfg.blockDim = dim3(G::THREADS);
  cfg.dynamicSmemBytes = G::SMEM_BYTES;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeClusterDimension;
  at[0].val.clusterDim.x = G::CLS; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
  cfg.attrs = at;
  cfg.numAttrs = 1;
  int n = 0;
  if (cudaOccupancyMaxActiveClusters(&n, lm_kernel<KS>, &cfg) != cudaSuccess) {
    cudaGetLastError();
    n = 0;
  }
  return n;
}

template <int KS>
static void launch_variant(const LmLayer* table, int nlayers, int rows, int nclusters, cudaStream_t st) {
  using G = Geo<KS>;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(G::CLS * nclusters);
  cfg.blockDim = dim3(G::THREADS);
  cfg.dynamicSmemBytes = G::SMEM_BYTES;
  cfg.stream = st;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeClusterDimension;
  at[0].val.clusterDim.x = G::CLS; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
  cfg.attrs = at;
  cfg.numAttrs = 1;
  cudaLaunchKernelEx(&cfg, lm_kernel<KS>, table, nlayers, rows, g_dbg_dev,
                     g_prof_on ? g_prof_dev : (unsigned long long*)nullptr, g_flags);
}
}  // namespace lm
}  // namespace dd

extern "C" int dd_program_run(const void* table_dev, int nlayers, int rows, dd_stream_t stream) {
  if (lm::g_max_clusters[1] == 0 && lm::g_max_clusters[2] == 0) {
    lm::g_max_clusters[1] = lm::setup_variant<1>();
    lm::g_max_clusters[2] = lm::setup_variant<2>();
    if (lm::g_ks < 0) {
      // default: clusters of 8.  Measured on B200 (profiles/r02_machine_geometry.txt): only 7
      // clusters of 16 CTAs are co-resident (112 SMs), so the 8 row tiles of the a1 rollout take
      // two rounds and large batches lose to 16 clusters of 8 (128 SMs); the split-K geometry
      // stays available (DD_MACHINE_KS=2 / dd_set_machine_ks) for parts that host 8 such clusters.
      const char* e = getenv("DD_MACHINE_KS");
      lm::g_ks = e != nullptr ? atoi(e) : 1;
    }
    if (lm::g_ks == 2 && lm::g_max_clusters[2] < 1) lm::g_ks = 1;
    if (lm::g_ks != 2) lm::g_ks = 1;
    if (lm::g_max_clusters[lm::g_ks] < 1) return fail("dd_program_run", "no cluster of the layer machine fits this device");
  }
  const int ks = lm::g_ks;
  const int ntiles = (rows + lm::BM - 1) / lm::BM;
  const int nclusters = ntiles < lm::g_max_clusters[ks] ? ntiles : lm::g_max_clusters[ks];
  const lm::LmLayer* table = static_cast<const lm::LmLayer*>(table_dev);
  if (ks == 2) lm::launch_variant<2>(table, nlayers, rows, nclusters, S(stream));
  else lm::launch_variant<1>(table, nlayers, rows, nclusters, S(stream));
  return launched("dd_program_run");
}

/* 1: clusters of 8 CTAs; 2: split-K clusters of 16.  Returns the previous choice (0 = unset). */
extern "C" int dd_set_machine_ks(int ks) {
  const int prev = lm::g_ks < 0 ? 0 : lm::g_ks;
  lm::g_ks = ks == 1 ? 1 : 2;
  if (lm::g_ks == 2 && lm::g_max_clusters[1] + lm::g_max_clusters[2] > 0 && lm::g_max_clusters[2] < 1) lm::g_ks = 1;
  return prev;
}

extern "C" int dd_program_debug(unsigned* out8) {
  if (lm::g_dbg_host == nullptr) return 0;
  for (int i = 0; i < 8; ++i) out8[i] = lm::g_dbg_host[i];
  return lm::g_dbg_host[0] == 0xDEADBEEFu ? 1 : 0;
}

/* profiling aids (not part of the product path): per-CTA wait-cycle counters and switches that
 * drop parts of the pipeline (results become garbage) */
extern "C" int dd_set_machine_debug(int flags) {
  const int prev = lm::g_flags;
  lm::g_flags = flags;
  return prev;
}
extern "C" int dd_machine_prof(int on, unsigned long long* out, int n) {
  if (lm::g_prof_dev == nullptr) {
    if (cudaMalloc(reinterpret_cast<void**>(&lm::g_prof_dev), (148 * 16 + 4 * lm::TRACE_N) * 8) != cudaSuccess) return -1;
    cudaMemset(lm::g_prof_dev, 0, (148 * 16 + 4 * lm::TRACE_N) * 8);
  }
  if (out != nullptr && n > 0) {
    const int cap = 148 * 16 + 4 * lm::TRACE_N;
    cudaMemcpy(out, lm::g_prof_dev, (size_t)(n < cap ? n : cap) * 8, cudaMemcpyDeviceToHost);
    cudaMemset(lm::g_prof_dev, 0, (size_t)cap * 8);
  }
  lm::g_prof_on = on;
  return 0;
}

extern "C" int dd_split_lo(const float* x, int ldx, float* lo, int ldlo, int rows, int cols,
                           dd_stream_t stream) {
  if (rows < 1 || cols < 1) return 0;
  lm::split_lo_kernel<<<grid_for((int64_t)rows * cols, 256), 256, 0, S(stream)>>>(x, ldx, lo, ldlo, rows, cols);
  return launched("dd_split_lo");
}

extern "C" int dd_transpose_lo(const float* p, float* pt, float* pt_lo, int njobs, const int* src,
                               const int* rows, const int* cols, const int* dst, const int* ldd,
                               const int* mode, dd_stream_t stream) {
  if (njobs < 1) return 0;
  if (njobs > DD_MAX_TRANSPOSE_JOBS) return fail("dd_transpose_lo", "too many matrices in one call");
  lm::TrJobs jobs;
  jobs.n = njobs;
  int tiles = 0;
  for (int j = 0; j < njobs; ++j) {
    jobs.src[j] = src[j]; jobs.rows[j] = rows[j]; jobs.cols[j] = cols[j]; jobs.dst[j] = dst[j];
    jobs.ldd[j] = ldd[j]; jobs.mode[j] = mode[j];
    jobs.tile0[j] = tiles;
    tiles += ((rows[j] + 31) / 32) * ((cols[j] + 31) / 32);
  }
  jobs.tile0[njobs] = tiles;
  lm::transpose_lo_kernel<<<tiles, 256, 0, S(stream)>>>(p, pt, pt_lo, jobs);
  return launched("dd_transpose_lo");
}
