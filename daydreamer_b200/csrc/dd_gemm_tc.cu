// tcgen05 (5th-gen tensor core) GEMM with fp32-class accuracy: 3xTF32.
//
//   C[M,N] = op(A) op(B) (+ bias) + beta*C,  fp32 in / fp32 out.
//
// Every fp32 operand x is split on the fly into hi = tf32(x) and lo = tf32(x-hi)
// and the product is accumulated as hi*hi + lo*hi + hi*lo in a TMEM fp32
// accumulator (the dropped lo*lo term is ~2^-22 relative).  This keeps the
// learner inside the 1e-4 parity bar of the reference's fp32 CPU path while
// running the Linear layers (reference nets.py:566-579) on tensor cores.
//
// Structure (one 128 x BN output tile per CTA, 288 threads):
//   warps 0-7  loaders: LDG.128 along the contiguous dim of each operand, split
//              into hi/lo, STS.128 into the *canonical no-swizzle UMMA layout*
//              (8x16-byte core matrices), fence.proxy.async, mbarrier arrive.
//              After the main loop the same warps run the epilogue
//              (tcgen05.ld 32x32b -> bias/beta -> global, or red.add for split-K).
//   warp 8     allocates TMEM, then one elected lane issues tcgen05.mma
//              kind::tf32 (M=128, N=BN, K=8), 3 per k-step, and tcgen05.commit
//              to recycle the smem stage / publish the accumulator.
// Operand majors: a row-major [rows, K] operand is K-major, a [K, rows] operand
// is MN-major; both are native for kind::tf32 so NN / NT / TN / TT need no
// transposes.  Requirements (checked by the host wrapper, which falls back to
// the SIMT kernel otherwise): M > 64, contiguous extents and leading
// dimensions multiples of 4 floats, 16-byte aligned bases.
#include "dd_common.cuh"

namespace dd {
int g_tc_debug = 0;
int g_tc_persistent = 1;
int g_tc_staged = 1;        // cp.async-staged operand loader for the 64-wide tile (r02: -1.3% step time, tests green)
namespace tc {

constexpr int BM = 128;
constexpr int BK = 32;
constexpr int LOADER_WARPS = 8;                 // two groups of 4 alternate k-blocks
#ifndef DD_TC_GROUP_WARPS
#define DD_TC_GROUP_WARPS 4
#endif
constexpr int GROUP_WARPS = DD_TC_GROUP_WARPS;   // loader groups = 8 / GROUP_WARPS; must be <= STAGES (r01: 4 groups with
                                                 // 3 stages alias the mbarrier parity and hang; with 4 stages no faster)
constexpr int GROUP_THREADS = GROUP_WARPS * 32;
constexpr int NGROUPS = 8 / GROUP_WARPS;
constexpr int LOADER_THREADS = LOADER_WARPS * 32;
constexpr int THREADS = LOADER_THREADS + 32;

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t done;
  do {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.b32 %0, 1, 0, p;\n\t}"
        : "=r"(done)
        : "r"(bar), "r"(parity)
        : "memory");
  } while (!done);
}
// Round-to-nearest (ties away) to tf32 with two ALU ops.  cvt.rna.tf32.f32 gives
// the same value but issues on the 16/clk/SM conversion pipe, which made the
// loaders -- not the tensor core -- the bottleneck (profiles/r01_tc_dissect.txt).
// (r01 experiment: electing one lane per warp for the mbarrier wait/arrive made every
// shape 10-20% SLOWER than letting all threads arrive; kept for reference, unused.)
__device__ __forceinline__ void warp_mbar_wait(uint32_t bar, uint32_t parity, int lane) {
  if (lane == 0) mbar_wait(bar, parity);
  __syncwarp();
}
__device__ __forceinline__ void warp_mbar_arrive(uint32_t bar, int lane) {
  __syncwarp();
  if (lane == 0) mbar_arrive(bar);
}

__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.b32 %0, 1, 0, p;\n\t}" : "=r"(pred));
  return pred != 0;
}

__device__ __forceinline__ float to_tf32(float x) {
  return __uint_as_float((__float_as_uint(x) + 0x1000u) & 0xFFFFE000u);
}

// smem matrix descriptor (cute::UMMA::SmemDescriptor bit layout: start>>4 [0,14),
// LBO>>4 [16,30), SBO>>4 [32,46), version=1 [46,48), layout type [61,64):
// 0 = no swizzle, 1 = SWIZZLE_128B_BASE32B)
__device__ __forceinline__ uint64_t make_desc(uint32_t addr, uint32_t lbo, uint32_t sbo,
                                              uint32_t layout_type) {
  return (uint64_t)((addr & 0x3FFFFu) >> 4) | ((uint64_t)((lbo >> 4) & 0x3FFFu) << 16) |
         ((uint64_t)((sbo >> 4) & 0x3FFFu) << 32) | (1ull << 46) |
         ((uint64_t)layout_type << 61);
}

// kind::tf32 instruction descriptor (cute::UMMA::InstrDescriptor): c=F32 [4,6),
// a=TF32 [7,10), b=TF32 [10,13), a_major [15], b_major [16], N>>3 [17,23), M>>4 [24,29)
__host__ __device__ constexpr uint32_t make_idesc(int a_mn_major, int b_mn_major, int n) {
  return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)a_mn_major << 15) |
         ((uint32_t)b_mn_major << 16) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);
}

__device__ __forceinline__ void mma_tf32(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc,
                                         uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, {%5, %5, %5, %5}, p;\n\t}"
      :
      : "r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate), "r"(0u)
      : "memory");
}

__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&v)[32]) {
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
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

// One operand tile of R rows (MN extent) x BK, hi and lo copies.
// MNMAJ == false: global element (r, k) at g[r*ld + k]  (K-major)
//   smem unit (r, kc=k/4) at (r%8)*16 + (r/8)*128 + kc*(R/8)*128   (SBO=128, LBO=R*16)
// MNMAJ == true : global element (r, k) at g[k*ld + r]  (MN-major).  For 32-bit
//   operands the only MN-major layout the tensor core accepts is
//   SWIZZLE_128B_BASE32B (cutlass sm100_common.inl:92): atoms of 32 r x 4 k
//   (4 rows of 128 B), the 32-byte chunk index XORed with the row index:
//   byte(r, k) = (k/4)*SBO + (r/32)*512 + (k%4)*128 + (((r%32)/8) ^ (k%4))*32 + (r%8)*4,
//   SBO = (R/32)*512, LBO = 512.
template <int R, bool MNMAJ>
struct TileMap {
  static constexpr int NBLK = (R / 8) * (BK / 16);   // 128 elements (32 float4 units) per block
  static constexpr int PER = NBLK / GROUP_WARPS;
  static_assert(NBLK % GROUP_WARPS == 0, "blocks per warp");
  // (row, k) of the float4 unit this lane handles in block `blk`, and its smem byte offset
  __device__ static __forceinline__ void at(int blk, int lane, int& r, int& k, uint32_t& off) {
    if (!MNMAJ) {
      const int rg = blk % (R / 8), half = blk / (R / 8);     // 8 rows x 4 k-units
      r = rg * 8 + (lane & 7);
      const int kc = half * 4 + (lane >> 3);
      k = kc * 4;
      off = (uint32_t)((r & 7) * 16 + (r >> 3) * 128 + kc * (R / 8) * 128);
    } else {
      const int ra = blk % (R / 32), kg = blk / (R / 32);     // one 32 r x 4 k atom
      const int kr = lane >> 3, unit = lane & 7;              // row in atom, 16-B unit in row
      k = kg * 4 + kr;
      r = ra * 32 + unit * 4;
      off = (uint32_t)(kg * (R / 32) * 512 + ra * 512 + kr * 128 +
                       (((unit >> 1) ^ kr) << 5) + ((unit & 1) << 4));
    }
  }
};

template <int R, bool MNMAJ>
__device__ __forceinline__ void tile_gload(const float* __restrict__ g, int ld, int r0, int rmax,
                                           int k0, int kmax, int ltid,
                                           float4 (&v)[TileMap<R, MNMAJ>::PER]) {
  const int lane = ltid & 31, warp = ltid >> 5;
#pragma unroll
  for (int it = 0; it < TileMap<R, MNMAJ>::PER; ++it) {
    int r, k;
    uint32_t off;
    TileMap<R, MNMAJ>::at(warp + it * GROUP_WARPS, lane, r, k, off);
    v[it] = make_float4(0.f, 0.f, 0.f, 0.f);
    const int gr = r0 + r, gk = k0 + k;
    if (gr < rmax && gk < kmax) {
      const float* p = MNMAJ ? g + (size_t)gk * ld + gr : g + (size_t)gr * ld + gk;
      v[it] = *reinterpret_cast<const float4*>(p);
    }
  }
}

template <int R, bool MNMAJ>
__device__ __forceinline__ void tile_split_store(const float4 (&v)[TileMap<R, MNMAJ>::PER],
                                                 char* s_hi, char* s_lo, int ltid) {
  const int lane = ltid & 31, warp = ltid >> 5;
#pragma unroll
  for (int it = 0; it < TileMap<R, MNMAJ>::PER; ++it) {
    int r, k;
    uint32_t off;
    TileMap<R, MNMAJ>::at(warp + it * GROUP_WARPS, lane, r, k, off);
    float4 hi, lo;
    hi.x = to_tf32(v[it].x); hi.y = to_tf32(v[it].y);
    hi.z = to_tf32(v[it].z); hi.w = to_tf32(v[it].w);
    // x - hi is exact in fp32; the tensor core ignores lo's low 13 mantissa bits
    lo.x = v[it].x - hi.x; lo.y = v[it].y - hi.y;
    lo.z = v[it].z - hi.z; lo.w = v[it].w - hi.w;
    *reinterpret_cast<float4*>(s_hi + off) = hi;
    *reinterpret_cast<float4*>(s_lo + off) = lo;
  }
}

// ---- staged loader (experimental, dd_set_gemm_tc_staged) ----------------------------------
// Same unit map as tile_gload, but the 16-byte units go global -> shared with cp.async into a
// per-thread slot of a raw ring instead of into registers, so a loader group can keep RD
// k-blocks of L2 latency in flight instead of one.  Each thread later re-reads exactly the
// units it copied itself (after cp.async.wait_group), so no cross-thread synchronisation is
// added.  Ring layout: [slot][unit][thread] x 16 B (consecutive threads -> consecutive banks).
template <int R, bool MNMAJ>
__device__ __forceinline__ void tile_cpasync(const float* __restrict__ g, int ld, int r0, int rmax,
                                             int k0, int kmax, int ltid, uint32_t raw_unit0) {
  const int lane = ltid & 31, warp = ltid >> 5;
#pragma unroll
  for (int it = 0; it < TileMap<R, MNMAJ>::PER; ++it) {
    int r, k;
    uint32_t off;
    TileMap<R, MNMAJ>::at(warp + it * GROUP_WARPS, lane, r, k, off);
    const int gr = r0 + r, gk = k0 + k;
    const bool in = gr < rmax && gk < kmax;
    const float* p = in ? (MNMAJ ? g + (size_t)gk * ld + gr : g + (size_t)gr * ld + gk) : g;
    const uint32_t dst = raw_unit0 + (uint32_t)it * GROUP_THREADS * 16;
    const uint32_t bytes = in ? 16u : 0u;               // 0 -> the 16 bytes are zero-filled
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(p), "r"(bytes)
                 : "memory");
  }
}

template <int R, bool MNMAJ>
__device__ __forceinline__ void tile_split_store_raw(const char* raw_unit0, char* s_hi, char* s_lo,
                                                     int ltid) {
  const int lane = ltid & 31, warp = ltid >> 5;
#pragma unroll
  for (int it = 0; it < TileMap<R, MNMAJ>::PER; ++it) {
    int r, k;
    uint32_t off;
    TileMap<R, MNMAJ>::at(warp + it * GROUP_WARPS, lane, r, k, off);
    const float4 v = *reinterpret_cast<const float4*>(raw_unit0 + (size_t)it * GROUP_THREADS * 16);
    float4 hi, lo;
    hi.x = to_tf32(v.x); hi.y = to_tf32(v.y); hi.z = to_tf32(v.z); hi.w = to_tf32(v.w);
    lo.x = v.x - hi.x; lo.y = v.y - hi.y; lo.z = v.z - hi.z; lo.w = v.w - hi.w;
    *reinterpret_cast<float4*>(s_hi + off) = hi;
    *reinterpret_cast<float4*>(s_lo + off) = lo;
  }
}

constexpr int RAW_DEPTH = 2;       // raw k-blocks in flight per loader group (staged loader)
constexpr int RAW_OFFSET = 1024;   // raw ring starts this far after the hi/lo stages (barriers live there)

template <int BN, int STAGES, bool TA, bool TB, bool STAGED = false>
__global__ void __launch_bounds__(THREADS, 1)
gemm_tc_kernel(const float* __restrict__ A, int lda, const float* __restrict__ B, int ldb,
               float* __restrict__ C, int ldc, int M, int N, int K, float beta,
               const float* __restrict__ bias, int klen, int atomic, int dbg) {
  pdl_prologue();
  // dbg (profiling only): bit 0 skips the loader's global loads / smem stores,
  // bit 1 skips the MMA issue; results are garbage, timing isolates each side.
  // A operand: TA ? MN-major ([K, M] in memory) : K-major.  B operand (n, k):
  // TB ? K-major ([N, K] in memory) : MN-major ([K, N] in memory).
  constexpr bool A_MN = TA, B_MN = !TB;
  constexpr int A_BYTES = BM * BK * 4, B_BYTES = BN * BK * 4;
  constexpr int STAGE_BYTES = 2 * A_BYTES + 2 * B_BYTES;
  extern __shared__ __align__(1024) char smem[];
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + STAGES * STAGE_BYTES);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * STAGES + 1);

  // warp index broadcast from lane 0: the role dispatch compiles as warp-uniform control flow
  const int tid = threadIdx.x, warp = __shfl_sync(0xffffffffu, tid >> 5, 0), lane = tid & 31;
  const int m0 = blockIdx.y * BM, n0 = blockIdx.x * BN;
  const int kbeg = blockIdx.z * klen;
  const int kend = min(K, kbeg + klen);
  const int nkb = kend > kbeg ? (kend - kbeg + BK - 1) / BK : 0;

  const uint32_t full0 = smem_u32(bars), empty0 = smem_u32(bars + STAGES);
  const uint32_t accum_bar = smem_u32(bars + 2 * STAGES);

  if (tid == 0) {
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(full0 + 8 * s, GROUP_THREADS);
      mbar_init(empty0 + 8 * s, 1);
    }
    mbar_init(accum_bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == LOADER_WARPS) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(
                     smem_u32(tmem_slot)),
                 "r"((uint32_t)(2 * BN))
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = *tmem_slot;

  if (warp < LOADER_WARPS) {
    // ===== producers: global -> hi/lo split -> canonical smem =================
    // Two loader groups (warps 0-3 / 4-7) take alternate k-blocks; each keeps the
    // global loads of its NEXT block in flight while it splits/stores the current
    // one, so ~4 k-blocks of L2 latency overlap with the tensor-core work.
    const int grp = warp / GROUP_WARPS, gtid = tid % GROUP_THREADS;
    if constexpr (STAGED) {
      static_assert(NGROUPS <= STAGES, "a producer may not run two mbarrier phases ahead");
      constexpr int UA = TileMap<BM, A_MN>::PER, UB = TileMap<BN, B_MN>::PER;
      constexpr uint32_t SLOT_BYTES = (uint32_t)(UA + UB) * GROUP_THREADS * 16;
      char* rawbase = smem + STAGES * STAGE_BYTES + RAW_OFFSET;
      auto slot_ptr = [&](int slot) { return rawbase + (size_t)slot * SLOT_BYTES + (size_t)gtid * 16; };
      auto issue = [&](int kbp, int slot) {
        const uint32_t u0 = smem_u32(slot_ptr(slot));
        tile_cpasync<BM, A_MN>(A, lda, m0, M, kbeg + kbp * BK, kend, gtid, u0);
        tile_cpasync<BN, B_MN>(B, ldb, n0, N, kbeg + kbp * BK, kend, gtid,
                               u0 + (uint32_t)UA * GROUP_THREADS * 16);
      };
#pragma unroll
      for (int d = 0; d < RAW_DEPTH; ++d) {             // one commit group per k-block, even if empty
        const int kbp = grp + NGROUPS * d;
        if (kbp < nkb) issue(kbp, grp * RAW_DEPTH + d);
        asm volatile("cp.async.commit_group;" ::: "memory");
      }
      int i = 0;
      for (int kb = grp; kb < nkb; kb += NGROUPS, ++i) {
        asm volatile("cp.async.wait_group %0;" ::"n"(RAW_DEPTH - 1) : "memory");
        const int slot = grp * RAW_DEPTH + (i % RAW_DEPTH);
        const int s = kb % STAGES;
        const uint32_t ph = (kb / STAGES) & 1;
        mbar_wait(empty0 + 8 * s, ph ^ 1);
        char* st = smem + s * STAGE_BYTES;
        const char* ru = slot_ptr(slot);
        tile_split_store_raw<BM, A_MN>(ru, st, st + A_BYTES, gtid);
        tile_split_store_raw<BN, B_MN>(ru + (size_t)UA * GROUP_THREADS * 16, st + 2 * A_BYTES,
                                       st + 2 * A_BYTES + B_BYTES, gtid);
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        mbar_arrive(full0 + 8 * s);
        const int kbn = kb + NGROUPS * RAW_DEPTH;        // refill the slot this thread just drained
        if (kbn < nkb) issue(kbn, slot);
        asm volatile("cp.async.commit_group;" ::: "memory");
      }
      asm volatile("cp.async.wait_group 0;" ::: "memory");
    }
    float4 ra[TileMap<BM, A_MN>::PER], rb[TileMap<BN, B_MN>::PER];
    if (!STAGED && grp < nkb && !(dbg & 1)) {
      tile_gload<BM, A_MN>(A, lda, m0, M, kbeg + grp * BK, kend, gtid, ra);
      tile_gload<BN, B_MN>(B, ldb, n0, N, kbeg + grp * BK, kend, gtid, rb);
    }
    for (int kb = grp; !STAGED && kb < nkb; kb += NGROUPS) {
      const int s = kb % STAGES;
      const uint32_t ph = (kb / STAGES) & 1;
      mbar_wait(empty0 + 8 * s, ph ^ 1);
      char* st = smem + s * STAGE_BYTES;
      if (!(dbg & 1)) {
        if (!(dbg & 8)) {
          tile_split_store<BM, A_MN>(ra, st, st + A_BYTES, gtid);
          tile_split_store<BN, B_MN>(rb, st + 2 * A_BYTES, st + 2 * A_BYTES + B_BYTES, gtid);
        } else {                                       // dbg 8: consume the loads, skip the stores
          float acc = 0.f;
#pragma unroll
          for (int i = 0; i < TileMap<BM, A_MN>::PER; ++i) acc += ra[i].x + ra[i].w;
#pragma unroll
          for (int i = 0; i < TileMap<BN, B_MN>::PER; ++i) acc += rb[i].x + rb[i].w;
          if (acc == 1.2345e-30f) *reinterpret_cast<float*>(st) = acc;
        }
      }
      if (kb + NGROUPS < nkb && !(dbg & 1) && !(dbg & 16)) {  // dbg 16: no global loads after the first
        const int k1 = kbeg + (kb + NGROUPS) * BK;
        tile_gload<BM, A_MN>(A, lda, m0, M, k1, kend, gtid, ra);
        tile_gload<BN, B_MN>(B, ldb, n0, N, k1, kend, gtid, rb);
      }
      if (!(dbg & 4)) asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // dbg 4
      mbar_arrive(full0 + 8 * s);
    }
    // ===== epilogue: TMEM -> registers -> global =============================
    if (nkb > 0) {
      mbar_wait(accum_bar, 0);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    }
    const int quad = warp & 3, chalf = warp >> 2;        // TMEM lane quadrant, column half
    const int m = m0 + quad * 32 + lane;
    const bool lead = blockIdx.z == 0;
    constexpr int CHUNKS = BN / 32;
    for (int ch = chalf; ch < CHUNKS; ch += 2) {
      uint32_t v[32];
      if (nkb > 0) {
        uint32_t w[32];
        const uint32_t ta0 = tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(ch * 32);
        tmem_ld32(ta0, v);
        tmem_ld32(ta0 + BN, w);
#pragma unroll
        for (int j = 0; j < 32; ++j)
          v[j] = __float_as_uint(__uint_as_float(v[j]) + __uint_as_float(w[j]));
      } else {
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] = 0u;
      }
      const int nb = n0 + ch * 32;
      const bool vec = !atomic && (ldc & 3) == 0 && nb + 32 <= N &&
                       (reinterpret_cast<uintptr_t>(C) & 15) == 0;
      if (m < M && vec) {
        float4* crow4 = reinterpret_cast<float4*>(C + (size_t)m * ldc + nb);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          float4 x = make_float4(__uint_as_float(v[4 * j]), __uint_as_float(v[4 * j + 1]),
                                 __uint_as_float(v[4 * j + 2]), __uint_as_float(v[4 * j + 3]));
          if (bias != nullptr) {
            const float4 b = *reinterpret_cast<const float4*>(bias + nb + 4 * j);
            x.x += b.x; x.y += b.y; x.z += b.z; x.w += b.w;
          }
          if (beta != 0.f) {
            const float4 c = crow4[j];
            x.x += beta * c.x; x.y += beta * c.y; x.z += beta * c.z; x.w += beta * c.w;
          }
          crow4[j] = x;
        }
      } else if (m < M) {
        float* crow = C + (size_t)m * ldc;
#pragma unroll
        for (int j = 0; j < 32; ++j) {
          const int n = n0 + ch * 32 + j;
          if (n < N) {
            float x = __uint_as_float(v[j]);
            if (bias != nullptr && lead) x += bias[n];
            if (atomic) atomicAdd(&crow[n], x);
            else crow[n] = (beta != 0.f) ? x + beta * crow[n] : x;
          }
        }
      }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  } else {
    // ===== MMA issuer =============================================================
    constexpr uint32_t IDESC = make_idesc(A_MN ? 1 : 0, B_MN ? 1 : 0, BN);
    // per-k-step (K=8) start-address advance and descriptor strides, in bytes
    // K-major (no swizzle): LBO = stride between the two 4-k chunks, SBO = 8-row group.
    // MN-major (128B_BASE32B): LBO = 32-r atom stride, SBO = 4-k group stride.
    constexpr uint32_t A_LBO = A_MN ? 512 : (BM / 8) * 128;
    constexpr uint32_t B_LBO = B_MN ? 512 : (BN / 8) * 128;
    constexpr uint32_t A_SBO = A_MN ? (BM / 32) * 512 : 128;
    constexpr uint32_t B_SBO = B_MN ? (BN / 32) * 512 : 128;
    constexpr uint32_t A_STEP = A_MN ? 2 * A_SBO : 2 * A_LBO;     // K=8 per instruction
    constexpr uint32_t B_STEP = B_MN ? 2 * B_SBO : 2 * B_LBO;
    constexpr uint32_t A_LT = A_MN ? 1u : 0u, B_LT = B_MN ? 1u : 0u;
    for (int kb = 0; kb < nkb; ++kb) {
      const int s = kb % STAGES;
      const uint32_t ph = (kb / STAGES) & 1;
      mbar_wait(full0 + 8 * s, ph);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      // all lanes walk the loop; only the MMA / commit instructions are predicated on the elected
      // lane, so their operands stay in uniform registers (with `if (lane == 0)` around a diverged
      // region every tcgen05.mma was wrapped in an ELECT + R2UR waterfall: ~80 cycles per MMA)
      if (elect_one()) {
        const uint32_t a_hi = smem_u32(smem + s * STAGE_BYTES), a_lo = a_hi + A_BYTES;
        const uint32_t b_hi = a_hi + 2 * A_BYTES, b_lo = b_hi + B_BYTES;
#pragma unroll
        for (int ks = 0; ks < BK / 8; ++ks) {
          const uint64_t dah = make_desc(a_hi + ks * A_STEP, A_LBO, A_SBO, A_LT);
          const uint64_t dal = make_desc(a_lo + ks * A_STEP, A_LBO, A_SBO, A_LT);
          const uint64_t dbh = make_desc(b_hi + ks * B_STEP, B_LBO, B_SBO, B_LT);
          const uint64_t dbl = make_desc(b_lo + ks * B_STEP, B_LBO, B_SBO, B_LT);
          // the small cross terms accumulate separately so their (truncating)
          // tensor-core adds do not touch the main accumulator
          if (!(dbg & 2)) {
            mma_tf32(tmem_base + BN, dal, dbh, IDESC, (kb | ks) != 0);
            mma_tf32(tmem_base + BN, dah, dbl, IDESC, 1u);
            mma_tf32(tmem_base, dah, dbh, IDESC, (kb | ks) != 0);
          }
        }
        // frees this smem stage once the MMAs above have read it
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::
                         "r"(empty0 + 8 * s)
                     : "memory");
        if (kb == nkb - 1)
          asm volatile(
              "tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(
                  accum_bar)
              : "memory");
      }
      __syncwarp();
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  }
  __syncthreads();
  if (warp == LOADER_WARPS) {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base),
                 "r"((uint32_t)(2 * BN))
                 : "memory");
  }
}

// ---------------------------------------------------------------------------
// Persistent variant for GEMMs with more output tiles than SMs (the M ~ 16K head
// and trajectory layers).  r01 measurement (profiles/r01_tc_gemm_dissect.txt):
// ~10 us of every one-tile CTA is fixed cost (launch, TMEM alloc, barrier setup
// and an epilogue nothing overlaps).  Here one CTA per SM walks a static tile
// schedule; the TMEM accumulator is double buffered (2 x (main + cross) = 4*BN
// = 512 columns) and 4 dedicated epilogue warps drain tile i while the loaders
// and the MMA warp already work on tile i+1.
//   warps 0-7 loaders (two alternating groups), warps 8-11 epilogue (TMEM lane
//   quadrant = warp % 4), warp 12 TMEM alloc + MMA issue.
// ---------------------------------------------------------------------------
constexpr int P_THREADS = 13 * 32;

template <int BN, int STAGES, bool TA, bool TB>
__global__ void __launch_bounds__(P_THREADS, 1)
gemm_tc_persistent_kernel(const float* __restrict__ A, int lda, const float* __restrict__ B, int ldb,
                          float* __restrict__ C, int ldc, int M, int N, int K, float beta,
                          const float* __restrict__ bias) {
  pdl_prologue();
  constexpr bool A_MN = TA, B_MN = !TB;
  constexpr int A_BYTES = BM * BK * 4, B_BYTES = BN * BK * 4;
  constexpr int STAGE_BYTES = 2 * A_BYTES + 2 * B_BYTES;
  extern __shared__ __align__(1024) char smem[];
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + STAGES * STAGE_BYTES);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * STAGES + 4);

  // warp index broadcast from lane 0: the role dispatch compiles as warp-uniform control flow
  const int tid = threadIdx.x, warp = __shfl_sync(0xffffffffu, tid >> 5, 0), lane = tid & 31;
  const int tiles_n = (N + BN - 1) / BN, tiles_m = (M + BM - 1) / BM;
  const int ntiles = tiles_m * tiles_n;
  const int nkb = (K + BK - 1) / BK;

  const uint32_t full0 = smem_u32(bars), empty0 = smem_u32(bars + STAGES);
  const uint32_t acc_full0 = smem_u32(bars + 2 * STAGES), acc_empty0 = smem_u32(bars + 2 * STAGES + 2);

  if (tid == 0) {
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(full0 + 8 * s, GROUP_THREADS);
      mbar_init(empty0 + 8 * s, 1);
    }
    for (int b = 0; b < 2; ++b) {
      mbar_init(acc_full0 + 8 * b, 1);
      mbar_init(acc_empty0 + 8 * b, 4);            // one arrive per epilogue warp
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 12) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(
                     smem_u32(tmem_slot)),
                 "r"((uint32_t)(4 * BN))
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = *tmem_slot;

  if (warp < LOADER_WARPS) {
    // ===== loaders: global k-block index `it` runs across tiles ==================
    const int grp = warp / GROUP_WARPS, gtid = tid % GROUP_THREADS;
    const int my_tiles = ntiles > (int)blockIdx.x ? (ntiles - 1 - (int)blockIdx.x) / (int)gridDim.x + 1 : 0;
    const long total = (long)my_tiles * nkb;
    float4 ra[TileMap<BM, A_MN>::PER], rb[TileMap<BN, B_MN>::PER];
    auto coords = [&](long it, int& m0, int& n0, int& k0) {
      const int ti = (int)(it / nkb);
      const int tile = (int)blockIdx.x + ti * (int)gridDim.x;
      m0 = (tile / tiles_n) * BM;
      n0 = (tile % tiles_n) * BN;
      k0 = (int)(it % nkb) * BK;
    };
    if (grp < total) {
      int m0, n0, k0;
      coords(grp, m0, n0, k0);
      tile_gload<BM, A_MN>(A, lda, m0, M, k0, K, gtid, ra);
      tile_gload<BN, B_MN>(B, ldb, n0, N, k0, K, gtid, rb);
    }
    for (long it = grp; it < total; it += NGROUPS) {
      const int s = (int)(it % STAGES);
      const uint32_t ph = (uint32_t)((it / STAGES) & 1);
      mbar_wait(empty0 + 8 * s, ph ^ 1);
      char* st = smem + s * STAGE_BYTES;
      tile_split_store<BM, A_MN>(ra, st, st + A_BYTES, gtid);
      tile_split_store<BN, B_MN>(rb, st + 2 * A_BYTES, st + 2 * A_BYTES + B_BYTES, gtid);
      if (it + NGROUPS < total) {
        int m0, n0, k0;
        coords(it + NGROUPS, m0, n0, k0);
        tile_gload<BM, A_MN>(A, lda, m0, M, k0, K, gtid, ra);
        tile_gload<BN, B_MN>(B, ldb, n0, N, k0, K, gtid, rb);
      }
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      mbar_arrive(full0 + 8 * s);
    }
  } else if (warp < 12) {
    // ===== epilogue warps: drain accumulator buffer (ti & 1) =========================
    const int quad = warp & 3;
    int ti = 0;
    for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++ti) {
      const int buf = ti & 1;
      const uint32_t ph = (uint32_t)((ti >> 1) & 1);
      mbar_wait(acc_full0 + 8 * buf, ph);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      const int m0 = (tile / tiles_n) * BM, n0 = (tile % tiles_n) * BN;
      const int m = m0 + quad * 32 + lane;
#pragma unroll 1
      for (int ch = 0; ch < BN / 32; ++ch) {
        uint32_t v[32], w[32];
        const uint32_t ta0 = tmem_base + ((uint32_t)(quad * 32) << 16) +
                             (uint32_t)(buf * 2 * BN + ch * 32);
        tmem_ld32(ta0, v);
        tmem_ld32(ta0 + BN, w);
#pragma unroll
        for (int j = 0; j < 32; ++j)
          v[j] = __float_as_uint(__uint_as_float(v[j]) + __uint_as_float(w[j]));
        const int nb = n0 + ch * 32;
        const bool vec = (ldc & 3) == 0 && nb + 32 <= N &&
                         (reinterpret_cast<uintptr_t>(C) & 15) == 0;
        if (m < M && vec) {
          float4* crow4 = reinterpret_cast<float4*>(C + (size_t)m * ldc + nb);
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            float4 x = make_float4(__uint_as_float(v[4 * j]), __uint_as_float(v[4 * j + 1]),
                                   __uint_as_float(v[4 * j + 2]), __uint_as_float(v[4 * j + 3]));
            if (bias != nullptr) {
              const float4 b = *reinterpret_cast<const float4*>(bias + nb + 4 * j);
              x.x += b.x; x.y += b.y; x.z += b.z; x.w += b.w;
            }
            if (beta != 0.f) {
              const float4 c = crow4[j];
              x.x += beta * c.x; x.y += beta * c.y; x.z += beta * c.z; x.w += beta * c.w;
            }
            crow4[j] = x;
          }
        } else if (m < M) {
          float* crow = C + (size_t)m * ldc;
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            const int n = nb + j;
            if (n < N) {
              float x = __uint_as_float(v[j]);
              if (bias != nullptr) x += bias[n];
              crow[n] = (beta != 0.f) ? x + beta * crow[n] : x;
            }
          }
        }
      }
      // this buffer may be overwritten by the MMAs of tile ti + 2
      asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
      __syncwarp();
      if (lane == 0) mbar_arrive(acc_empty0 + 8 * buf);
    }
  } else {
    // ===== MMA issuer ==================================================================
    constexpr uint32_t IDESC = make_idesc(A_MN ? 1 : 0, B_MN ? 1 : 0, BN);
    constexpr uint32_t A_LBO = A_MN ? 512 : (BM / 8) * 128;
    constexpr uint32_t B_LBO = B_MN ? 512 : (BN / 8) * 128;
    constexpr uint32_t A_SBO = A_MN ? (BM / 32) * 512 : 128;
    constexpr uint32_t B_SBO = B_MN ? (BN / 32) * 512 : 128;
    constexpr uint32_t A_STEP = A_MN ? 2 * A_SBO : 2 * A_LBO;
    constexpr uint32_t B_STEP = B_MN ? 2 * B_SBO : 2 * B_LBO;
    constexpr uint32_t A_LT = A_MN ? 1u : 0u, B_LT = B_MN ? 1u : 0u;
    long it = 0;
    int ti = 0;
    for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++ti) {
      const int buf = ti & 1;
      const uint32_t aph = (uint32_t)((ti >> 1) & 1);
      mbar_wait(acc_empty0 + 8 * buf, aph ^ 1);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      const uint32_t tacc = tmem_base + (uint32_t)(buf * 2 * BN);
      for (int kb = 0; kb < nkb; ++kb, ++it) {
        const int s = (int)(it % STAGES);
        const uint32_t ph = (uint32_t)((it / STAGES) & 1);
        mbar_wait(full0 + 8 * s, ph);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        if (elect_one()) {
          const uint32_t a_hi = smem_u32(smem + s * STAGE_BYTES), a_lo = a_hi + A_BYTES;
          const uint32_t b_hi = a_hi + 2 * A_BYTES, b_lo = b_hi + B_BYTES;
#pragma unroll
          for (int ks = 0; ks < BK / 8; ++ks) {
            const uint64_t dah = make_desc(a_hi + ks * A_STEP, A_LBO, A_SBO, A_LT);
            const uint64_t dal = make_desc(a_lo + ks * A_STEP, A_LBO, A_SBO, A_LT);
            const uint64_t dbh = make_desc(b_hi + ks * B_STEP, B_LBO, B_SBO, B_LT);
            const uint64_t dbl = make_desc(b_lo + ks * B_STEP, B_LBO, B_SBO, B_LT);
            mma_tf32(tacc + BN, dal, dbh, IDESC, (kb | ks) != 0);
            mma_tf32(tacc + BN, dah, dbl, IDESC, 1u);
            mma_tf32(tacc, dah, dbh, IDESC, (kb | ks) != 0);
          }
          asm volatile(
              "tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(
                  empty0 + 8 * s)
              : "memory");
          if (kb == nkb - 1)
            asm volatile(
                "tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(
                    acc_full0 + 8 * buf)
                : "memory");
        }
        __syncwarp();
      }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  }
  __syncthreads();
  if (warp == 12) {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base),
                 "r"((uint32_t)(4 * BN))
                 : "memory");
  }
}

template <int BN, int STAGES>
static int launch_persistent(const float* A, int lda, int ta, const float* B, int ldb, int tb,
                             float* C, int ldc, int M, int N, int K, float beta, const float* bias,
                             cudaStream_t st) {
  constexpr int STAGE_BYTES = 2 * BM * BK * 4 + 2 * BN * BK * 4;
  constexpr int SMEM = STAGES * STAGE_BYTES + (2 * STAGES + 4) * 8 + 16;
  const int ntiles = ((M + BM - 1) / BM) * ((N + BN - 1) / BN);
  const int grid = ntiles < 148 ? ntiles : 148;
#define DD_TCP(TA_, TB_)                                                                       \
  do {                                                                                         \
    static bool attr_set = false;                                                              \
    auto kern = gemm_tc_persistent_kernel<BN, STAGES, TA_, TB_>;                               \
    if (!attr_set) {                                                                           \
      cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM);           \
      attr_set = true;                                                                         \
    }                                                                                          \
    launch_k(kern, grid, P_THREADS, SMEM, st, A, lda, B, ldb, C, ldc, M, N, K, beta, bias);          \
  } while (0)
  if (!ta && !tb) DD_TCP(false, false);
  else if (ta && !tb) DD_TCP(true, false);
  else if (!ta && tb) DD_TCP(false, true);
  else DD_TCP(true, true);
#undef DD_TCP
  return launched("dd_gemm_tc_persistent");
}

template <int BN, int STAGES>
static int launch(const float* A, int lda, int ta, const float* B, int ldb, int tb, float* C,
                  int ldc, int M, int N, int K, float beta, const float* bias, int splits,
                  cudaStream_t st) {
  constexpr int STAGE_BYTES = 2 * BM * BK * 4 + 2 * BN * BK * 4;
  constexpr int SMEM = STAGES * STAGE_BYTES + (2 * STAGES + 1) * 8 + 16;
  dim3 grid((N + BN - 1) / BN, (M + BM - 1) / BM, splits);
  int klen = (K + splits - 1) / splits;
  klen = (klen + BK - 1) / BK * BK;
  const int atomic = splits > 1 ? 1 : 0;
#define DD_TC(TA_, TB_)                                                                        \
  do {                                                                                         \
    static bool attr_set = false;                                                              \
    auto kern = gemm_tc_kernel<BN, STAGES, TA_, TB_>;                                          \
    if (!attr_set) {                                                                           \
      cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM);           \
      attr_set = true;                                                                         \
    }                                                                                          \
    launch_k(kern, grid, THREADS, SMEM, st, A, lda, B, ldb, C, ldc, M, N, K, beta, bias, klen, atomic, \
                                      g_tc_debug);                                             \
  } while (0)
  if (!ta && !tb) DD_TC(false, false);
  else if (ta && !tb) DD_TC(true, false);
  else if (!ta && tb) DD_TC(false, true);
  else DD_TC(true, true);
#undef DD_TC
  return launched("dd_gemm_tc");
}

// Staged-loader launch (BN = 64 only: two hi/lo stages + a raw ring of NGROUPS * RAW_DEPTH
// k-blocks fit the 227 KB of one SM; the 128-wide tile has no room for the ring).
static int launch_staged(const float* A, int lda, int ta, const float* B, int ldb, int tb, float* C,
                         int ldc, int M, int N, int K, float beta, const float* bias, int splits,
                         cudaStream_t st) {
  constexpr int BN = 64, STAGES = 2;
  constexpr int STAGE_BYTES = 2 * BM * BK * 4 + 2 * BN * BK * 4;
  constexpr int UNITS = (BM / 8) * (BK / 16) / GROUP_WARPS + (BN / 8) * (BK / 16) / GROUP_WARPS;
  constexpr int SMEM = STAGES * STAGE_BYTES + RAW_OFFSET + NGROUPS * RAW_DEPTH * UNITS * GROUP_THREADS * 16;
  static_assert(SMEM <= 227 * 1024, "staged loader ring does not fit");
  dim3 grid((N + BN - 1) / BN, (M + BM - 1) / BM, splits);
  int klen = (K + splits - 1) / splits;
  klen = (klen + BK - 1) / BK * BK;
  const int atomic = splits > 1 ? 1 : 0;
#define DD_TCS(TA_, TB_)                                                                       \
  do {                                                                                         \
    static bool attr_set = false;                                                              \
    auto kern = gemm_tc_kernel<BN, STAGES, TA_, TB_, true>;                                    \
    if (!attr_set) {                                                                           \
      cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM);           \
      attr_set = true;                                                                         \
    }                                                                                          \
    launch_k(kern, grid, THREADS, SMEM, st, A, lda, B, ldb, C, ldc, M, N, K, beta, bias, klen, atomic, 0); \
  } while (0)
  if (!ta && !tb) DD_TCS(false, false);
  else if (ta && !tb) DD_TCS(true, false);
  else if (!ta && tb) DD_TCS(false, true);
  else DD_TCS(true, true);
#undef DD_TCS
  return launched("dd_gemm_tc_staged");
}

}  // namespace tc

// Returns 1 if the tensor-core path took the call, 0 if the shape/alignment is
// not eligible (caller falls back to the SIMT kernel), <0 on launch error.
int gemm_tc_try(const float* A, int lda, int ta, const float* B, int ldb, int tb, float* C, int ldc,
                int M, int N, int K, float beta, const float* bias, cudaStream_t st) {
  if (M <= 64 || N < 32 || K < 32) return 0;
  auto al16 = [](const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; };
  if (!al16(A) || !al16(B) || (lda & 3) || (ldb & 3)) return 0;
  // contiguous extents must be whole float4 units
  if ((ta ? M : K) & 3) return 0;
  if ((tb ? K : N) & 3) return 0;
  const int kNumSm = 148;
  const int tiles128 = ((M + 127) / 128) * ((N + 127) / 128);
  const int tiles64 = ((M + 127) / 128) * ((N + 63) / 64);
  auto pick_splits = [&](int ctas) {
    if (beta != 1.f || bias != nullptr) return 1;
    // very long reductions (conv weight gradients: K = images x pixels) are cut into
    // <= 16K-long pieces so the truncating TMEM accumulation stays at fp32-level error
    const int by_k = K > 32768 ? (K + 16383) / 16384 : 1;
    if (ctas >= kNumSm) return by_k > 256 ? 256 : by_k;
    int s = (kNumSm + ctas - 1) / ctas;
    const int maxs = K / (tc::BK * 4) > 1 ? K / (tc::BK * 4) : 1;
    if (s > maxs) s = maxs;
    s = s > 32 ? 32 : (s < 1 ? 1 : s);
    return s > by_k ? s : (by_k > 256 ? 256 : by_k);
  };
  int rc;
  if (tiles128 > kNumSm + kNumSm / 2 && g_tc_persistent && K >= tc::BK &&
      pick_splits(tiles128) == 1) {
    rc = tc::launch_persistent<128, 3>(A, lda, ta, B, ldb, tb, C, ldc, M, N, K, beta, bias, st);
  } else if (tiles128 >= kNumSm / 2) {
    rc = tc::launch<128, 3>(A, lda, ta, B, ldb, tb, C, ldc, M, N, K, beta, bias,
                            pick_splits(tiles128), st);
  } else if (g_tc_staged) {
    rc = tc::launch_staged(A, lda, ta, B, ldb, tb, C, ldc, M, N, K, beta, bias,
                           pick_splits(tiles64), st);
  } else {
    rc = tc::launch<64, 4>(A, lda, ta, B, ldb, tb, C, ldc, M, N, K, beta, bias,
                           pick_splits(tiles64), st);
  }
  return rc == 0 ? 1 : rc;
}

}  // namespace dd

extern "C" int dd_set_gemm_tc_persistent(int on) {
  const int prev = dd::g_tc_persistent;
  dd::g_tc_persistent = on ? 1 : 0;
  return prev;
}

extern "C" int dd_set_gemm_tc_staged(int on) {
  const int prev = dd::g_tc_staged;
  dd::g_tc_staged = on ? 1 : 0;
  return prev;
}

extern "C" int dd_set_gemm_tc_debug(int bits) {
  const int prev = dd::g_tc_debug;
  dd::g_tc_debug = bits;
  return prev;
}
