// lip_tc_dev.cuh — device-side building blocks shared by the tcgen05 kernels (lip_gemm_tc.cu, lip_conv_tc.cu):
// tile constants, mbarrier / TMA / tcgen05 PTX wrappers, UMMA descriptors, kernel parameters and the fused tile epilogue.
// Everything lives in an anonymous namespace: each translation unit gets its own copy.
#pragma once
#include <cuda.h>

#include "lip_common.cuh"

namespace lip {

namespace {

constexpr int TBM = 128;       // CTA tile rows (UMMA M)
constexpr int TBK = 32;        // fp32 elements per k-block = 128 bytes = one swizzle row
constexpr int UMMA_K = 8;      // tf32
constexpr int TC_THREADS = 384;   // 4 control warps + 8 drain/epilogue warps
// fp32 accumulation in TMEM rounds toward zero, a bias that grows with the number of accumulation steps.
// Two measures keep 3xTF32 inside the 1e-5 budget: the two cross terms (2^-11 smaller) accumulate in their own
// TMEM tile so they never truncate the main sum, and no TMEM accumulator lives longer than KC k-blocks
// (32 tf32 MMAs) before it is folded into the register accumulators in round-to-nearest.
constexpr int KC = 8;          // k-blocks per TMEM chunk (256 fp32 of K)
constexpr int STG_LD = 33;     // padded row of the epilogue staging tile

// ------------------------------------------------------------------------------------------------ PTX wrappers
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(bar), "r"(parity)
      : "memory");
  return ok != 0;
}
// bounded wait: a broken pipeline traps (-> cudaErrorLaunchFailure) instead of hanging the device
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  for (uint32_t it = 0; it < (1u << 26); ++it) {
    if (mbar_try_wait(bar, parity)) return;
  }
  __trap();
}

// One lane of a converged warp.  The role loops below run warp-uniformly and gate only the asynchronous issue with
// this predicate: operands then live in uniform registers and each UTCHMMA / UTMALDG is a single instruction
// (under `if (lane == 0)` the compiler wraps every one in a VOTEU/ELECT/R2UR.BROADCAST divergence loop, which
// costs more than the 64-cycle tf32 MMA itself).
__device__ __forceinline__ bool elect_one() {
  uint32_t pred = 0;
  asm volatile(
      "{\n\t.reg .b32 rx;\n\t.reg .pred px;\n\t"
      "elect.sync rx|px, %1;\n\t"
      "@px mov.s32 %0, 1;\n\t}"
      : "+r"(pred)
      : "r"(0xffffffffu));
  return pred != 0;
}

__device__ __forceinline__ void tma_load_3d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1, int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
      ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}
__device__ __forceinline__ void tma_load_3d_mc(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1, int c2,
                                               uint16_t mask) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster"
      " [%0], [%1, {%3, %4, %5}], [%2], %6;"
      ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "h"(mask)
      : "memory");
}
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ uint32_t cluster_id_x() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%clusterid.x;" : "=r"(r));
  return r;
}
__device__ __forceinline__ uint32_t num_clusters_x() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%nclusterid.x;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accum) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accum)
      : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}

__device__ __forceinline__ void umma_commit_mc(uint32_t bar, uint16_t mask) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
               ::"r"(bar), "h"(mask) : "memory");
}

// ---- cta_group::2 (CTA pair) forms ----------------------------------------------------------------------------
constexpr uint32_t PEER_BIT_MASK = 0xFEFFFFFFu;   // clears the pair-rank bit of a shared::cluster address -> even CTA
__device__ __forceinline__ void tma_load_3d_2sm(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1, int c2) {
  // executed by both CTAs of the pair: data lands in the executing CTA's smem, bytes are credited to the LEADER's barrier
  asm volatile(
      "cp.async.bulk.tensor.3d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
      ::"r"(dst), "l"(map), "r"(bar & PEER_BIT_MASK), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}
__device__ __forceinline__ void umma_tf32_2sm(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accum) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accum)
      : "memory");
}
__device__ __forceinline__ void umma_commit_2sm(uint32_t bar, uint16_t mask) {
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
               ::"r"(bar), "h"(mask) : "memory");
}
__device__ __forceinline__ void mbar_arrive_remote(uint32_t bar, uint32_t cta) {
  asm volatile(
      "{\n\t.reg .b32 ra;\n\t"
      "mapa.shared::cluster.u32 ra, %0, %1;\n\t"
      "mbarrier.arrive.shared::cluster.b64 _, [ra];\n\t}"
      ::"r"(bar), "r"(cta) : "memory");
}

__device__ __forceinline__ void tmem_ld32(uint32_t taddr, float (&v)[32]) {
  uint32_t r[32];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
}

__device__ __forceinline__ float tf32_rna(float x) {
  uint32_t r;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
  return __uint_as_float(r);
}

// ---- UMMA descriptors (cute/arch/mma_sm100_desc.hpp bit layout) ---------------------------------------------------
// shared-memory matrix descriptor, sm100 version field = 1.
// layout_type 2 = SWIZZLE_128B (16-byte swizzle atoms; K-major operands),
//             1 = SWIZZLE_128B_BASE32B (32-byte swizzle atoms, 4-row K groups): the ONLY layout tcgen05 accepts
//                 for MN-major 32-bit (tf32) operands (cutlass sm100_common.inl:92).
__device__ __forceinline__ uint64_t make_smem_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes,
                                                   uint32_t layout_type) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr & 0x3FFFF) >> 4);             // [0,14)  start address >> 4
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;    // [16,30) leading byte offset >> 4
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;    // [32,46) stride byte offset >> 4
  d |= (uint64_t)1 << 46;                              // [46,48) version = 1 (Blackwell)
  d |= (uint64_t)layout_type << 61;                    // [61,64) layout type
  return d;
}

__host__ __device__ constexpr uint32_t make_idesc(int M, int N, bool a_mn, bool b_mn) {
  return (1u << 4)                        // c_format = F32
         | (2u << 7)                      // a_format = TF32
         | (2u << 10)                     // b_format = TF32
         | ((a_mn ? 1u : 0u) << 15)       // a_major: 0 = K, 1 = MN
         | ((b_mn ? 1u : 0u) << 16)       // b_major
         | ((uint32_t)(N >> 3) << 17)     // n_dim
         | ((uint32_t)(M >> 4) << 24);    // m_dim
}

struct TcParams {
  int M, N, K1, K2, batch;
  int kc, merge;   // experiment knobs: k-blocks per TMEM chunk; cross terms share the main accumulator
  int dbg;   // microbenchmark knobs: 1 = skip epilogue global traffic, 2 = issue only the hi*hi MMA, 4 = skip TMA loads
  int a1_batched, b1_batched, a2_batched, b2_batched;
  float* C; float* C_lo;
  long long c_sz, c_sm;
  float scale;
  const float* bias; long long bias_sz;
  const float* mask; long long mask_sm;
  const float* add; long long add_sz; float add_scale;
  float* colsum; long long colsum_sz, colsum_ld;
  const int* b1_lo_nz;   // device flag: 0 -> B1's lo operand is identically zero (skip its loads and MMAs)
};


// ---- tile epilogue shared by both kernels ---------------------------------------------------------------------------
// A drain warp owns 32 tile rows (one TMEM lane quarter; lane <-> row) x 64 columns in `acc`.  Each 32x32 block is
// transposed through the warp's padded staging tile so that lane <-> column for the global accesses (128-byte
// coalesced rows), then  x = (scale * acc + bias[n]) * mask[m][n] + add_scale * add[m][n]  is stored as fp32 or as a
// TF32 (hi, lo) pair.  The row loops use hoisted base pointers and 8-row batches (loads first): the per-element
// instruction count, not memory bandwidth, was the cost of the first version of this epilogue (ncu: `no_inst`).
template <bool HAS_LO, int RB>   // RB: rows per load batch (all mask / add loads of a batch are in flight together)
__device__ __forceinline__ float store_block_rows(const TcParams& p, const float* __restrict__ sp, float* __restrict__ cp,
                                                  float* __restrict__ lp, const float* __restrict__ mp,
                                                  const float* __restrict__ ap, float bv, int nrows) {
  const float sc = p.scale, asc = p.add_scale;
  const long long cs = p.c_sm, ms = p.mask_sm;
  float colsum = 0.f;
  int r0 = 0;
  for (; r0 + RB <= nrows; r0 += RB) {
    float mv[RB], av[RB];
#pragma unroll
    for (int r = 0; r < RB; ++r) {
      mv[r] = mp ? __ldg(mp + (r0 + r) * ms) : 1.f;
      av[r] = ap ? __ldg(ap + (r0 + r) * cs) : 0.f;
    }
#pragma unroll
    for (int r = 0; r < RB; ++r) {
      const float x = fmaf(asc, av[r], fmaf(sc, sp[(r0 + r) * STG_LD], bv) * mv[r]);
      colsum += x;
      if (HAS_LO) {
        const float hh = tf32_rna(x);
        cp[(r0 + r) * cs] = hh;
        lp[(r0 + r) * cs] = tf32_rna(x - hh);
      } else {
        cp[(r0 + r) * cs] = x;
      }
    }
  }
  for (; r0 < nrows; ++r0) {
    const float mv = mp ? __ldg(mp + r0 * ms) : 1.f;
    const float av = ap ? __ldg(ap + r0 * cs) : 0.f;
    const float x = fmaf(asc, av, fmaf(sc, sp[r0 * STG_LD], bv) * mv);
    colsum += x;
    if (HAS_LO) {
      const float hh = tf32_rna(x);
      cp[r0 * cs] = hh;
      lp[r0 * cs] = tf32_rna(x - hh);
    } else {
      cp[r0 * cs] = x;
    }
  }
  return colsum;
}

// acc: the warp's 32 x (32 * NBLK) block (row = lane), m0/n0: tile origin, q: lane quarter, h: which column group
template <int NBLK, int RB>
__device__ __forceinline__ void tile_epilogue(const TcParams& p, float (&acc)[32 * NBLK], float* stg, int m0, int n0, int z,
                                              int q, int h, int lane) {
  const int mrow0 = m0 + q * 32;
  int nrows = p.M - mrow0;
  nrows = nrows > 32 ? 32 : nrows;
  const long long zc = (long long)z * p.c_sz + (long long)mrow0 * p.c_sm;
#pragma unroll
  for (int cc = 0; cc < NBLK; ++cc) {
    const int n = n0 + h * (32 * NBLK) + cc * 32 + lane;
#pragma unroll
    for (int i = 0; i < 32; ++i) stg[lane * STG_LD + i] = acc[cc * 32 + i];
    __syncwarp();
    if (n < p.N && !(p.dbg & 1)) {
      float csum = 0.f;
      if (nrows > 0) {
        const float bv = p.bias ? __ldg(p.bias + (long long)z * p.bias_sz + n) : 0.f;
        float* cp = p.C + zc + n;
        const float* mp = p.mask ? p.mask + (long long)mrow0 * p.mask_sm + n : nullptr;
        const float* ap = p.add ? p.add + (long long)z * p.add_sz + (long long)mrow0 * p.c_sm + n : nullptr;
        if (p.C_lo) csum = store_block_rows<true, RB>(p, stg + lane, cp, p.C_lo + zc + n, mp, ap, bv, nrows);
        else csum = store_block_rows<false, RB>(p, stg + lane, cp, nullptr, mp, ap, bv, nrows);
      }
      if (p.colsum) p.colsum[(long long)z * p.colsum_sz + (long long)(mrow0 >> 5) * p.colsum_ld + n] = csum;
    }
    __syncwarp();
  }
}

}  // namespace

// host-side tensor-map helpers (lip_gemm_tc.cu)
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeTiledFn tc_get_encode();
// kmajor: element (row r, k) at ptr + r*ld + k;  box {32 k, box_rows rows, 1}
// mn-major: element (k, col c) at ptr + k*ld + c; box {32 cols, 32 k, 1} (one box per 32-column chunk)
int tc_make_map(CUtensorMap* map, const float* ptr, bool kmajor, int64_t rows_or_cols, int64_t K, int64_t ld, int64_t sz,
                int64_t batch, int box_rows);

}  // namespace lip
