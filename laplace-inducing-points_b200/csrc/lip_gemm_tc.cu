// lip_gemm_tc.cu — tcgen05 (5th-gen tensor core) batched GEMM with 3xTF32 fp32 emulation for sm_100a.
//
// The hot GEMMs of the GGN-vector product (see lip_model.cu) in fp32-faithful arithmetic on the tensor cores:
//   D = A_hi*B_hi + A_hi*B_lo + A_lo*B_hi            (x = hi + lo, hi = tf32(x), lo = tf32(x - hi))
// accumulated in fp32 in TMEM.  Operands are pre-split (hi, lo) fp32 arrays in HBM (weights / activations are
// split once at bind time, intermediates are split by the producing kernel's epilogue, probe blocks by
// tf32_split), staged by TMA into 128B-swizzled shared-memory tiles, and consumed by tcgen05.mma.kind::tf32
// issued by one thread.  Three operand-major combinations cover the JVP, weight-gradient and delta-backprop
// GEMMs without any transposition pass:
//   JVP    : A K-major  (activations [M,K]),  B MN-major (tangent weights [K,N])      (+ a second A/B pair)
//   WGRAD  : A MN-major (activations^T),      B MN-major (deltas [K,N])
//   DGRAD  : A K-major  (deltas [M,K]),       B K-major  (weights [N,K])
// Three kernels share the PTX wrappers, the tile epilogue and the host-side tensor-map code:
//   gemm_tc_kernel   128 x 128 tiles, one CTA per SM                         (ragged tails, small problems)
//   gemm_tc2_kernel  256 x 128 tiles on CTA pairs (cta_group::2)             (JVP GEMMs with N = 128)
//   gemm_tc2w_kernel 256 x 256 tiles on CTA pairs, setmaxnreg re-balancing    (everything that fills them; see its header)
// Persistent CTAs walk a static tile schedule.  Warp roles (384 threads): warp 0 = TMA producer, warp 1 = MMA issuer,
// warp 2 = TMEM allocator, warps 4-11 = accumulator drain + epilogue.  fp32 accumulation in TMEM truncates, so the K loop is
// cut into chunks of KC k-blocks whose TMEM accumulators the drain warps fold, round-to-nearest, into fp32 REGISTER
// accumulators (two alternating TMEM buffers in the 128-wide kernels); the cross terms have their own TMEM tile.
// The tile epilogue transposes through shared memory so that mask/add reads and the stores are 128-byte coalesced.
// All mbarrier waits are bounded: a deadlock traps instead of hanging the GPU.
// LIP_TC_KC / LIP_TC_MERGE are experiment knobs (tools/tc_accuracy_exp.py -> profiles/r01_tc_accuracy*.txt).
#include <cuda.h>

#include <mutex>
#include <vector>
#include <stdlib.h>

#include "lip_tc_dev.cuh"

namespace lip {

namespace {

static int g_tc_dbg = 0;      // set only by lip_bench_tc_gemm
static int g_tc_force2 = -1;  // -1: environment / default, 0: 1-CTA kernel, 1: 2-CTA kernel


template <int BN>
struct SmemLayout {
  static constexpr int A_TILE = TBM * TBK * 4;   // 16 KB
  static constexpr int B_TILE = BN * TBK * 4;
  static constexpr int STAGE = 2 * A_TILE + 2 * B_TILE;
  static constexpr int STAGES = (BN <= 128) ? 3 : 2;
  static constexpr int STAGING = 8 * 32 * STG_LD * 4;   // one 32x32 (padded) transpose tile per drain warp
  static constexpr int BYTES = STAGES * STAGE + STAGING + 1024 /*align slack*/ + 256 /*barriers*/;
};

// One operand tile load (ROWS = TBM or BN rows of the tile's non-contraction index).
//   K-major : 3D map {K, rows, batch},  box {32 k, ROWS (or ROWS/2) rows, 1}: smem [row][32 k] (128 B rows)
//   MN-major: 3D map {cols, K, batch},  ROWS/32 boxes {32 cols, 32 k, 1}, 4 KB each: smem [chunk][k][32 cols]
// Out-of-range rows / columns / k are zero-filled by TMA.
// CL == 4 (2x2 cluster): this CTA fetches only half `half` of the tile and multicasts it to the CTAs in `mask`
// (the partner fetches the other half), so every tile is read from L2 once per CTA pair.
template <bool KMAJOR, int ROWS, int CL>
__device__ __forceinline__ void load_operand(uint32_t dst, const CUtensorMap* map, uint32_t bar, int k0, int row0, int z,
                                             int half, uint16_t mask) {
  if (CL == 1) {
    if (KMAJOR) {
      tma_load_3d(dst, map, bar, k0, row0, z);
    } else {
#pragma unroll
      for (int c = 0; c < ROWS / 32; ++c) tma_load_3d(dst + c * (TBK * 128), map, bar, row0 + 32 * c, k0, z);
    }
  } else {
    if (KMAJOR) {
      tma_load_3d_mc(dst + half * (ROWS / 2) * 128, map, bar, k0, row0 + half * (ROWS / 2), z, mask);
    } else {
#pragma unroll
      for (int c = 0; c < ROWS / 64; ++c) {
        const int ch = half * (ROWS / 64) + c;
        tma_load_3d_mc(dst + ch * (TBK * 128), map, bar, row0 + 32 * ch, k0, z, mask);
      }
    }
  }
}

template <int BN, bool A_K, bool B_K, int CL>
__global__ void __launch_bounds__(TC_THREADS, 1)
gemm_tc_kernel(const __grid_constant__ CUtensorMap mA1h, const __grid_constant__ CUtensorMap mA1l,
               const __grid_constant__ CUtensorMap mB1h, const __grid_constant__ CUtensorMap mB1l,
               const __grid_constant__ CUtensorMap mA2h, const __grid_constant__ CUtensorMap mA2l,
               const __grid_constant__ CUtensorMap mB2h, const __grid_constant__ CUtensorMap mB2l, TcParams p) {
  using SL = SmemLayout<BN>;
  static_assert(BN == 128, "register accumulators are sized for BN = 128");
  constexpr int TMEM_COLS = 512;          // 2 buffers x (cross-term tile + main tile) x BN columns
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t stg_base = smem_base + SL::STAGES * SL::STAGE;
  const uint32_t bar_base = stg_base + SL::STAGING;
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (SL::STAGES + s); };
  auto tfull_bar = [&](int b) { return bar_base + 8u * (2 * SL::STAGES + b); };
  auto tempty_bar = [&](int b) { return bar_base + 8u * (2 * SL::STAGES + 2 + b); };
  const uint32_t tmem_slot = bar_base + 8u * (2 * SL::STAGES + 4);

  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);   // provably warp-uniform
  const int lane = threadIdx.x & 31;
  const int nk1 = (p.K1 + TBK - 1) / TBK, nk2 = (p.K2 + TBK - 1) / TBK;
  const int nk = nk1 + nk2;
  const bool skip_b1lo = p.b1_lo_nz != nullptr && *reinterpret_cast<const volatile int*>(p.b1_lo_nz) == 0;
  const int KCr = p.kc;
  const int nchunks = (nk + KCr - 1) / KCr;
  const int mt = (p.M + TBM - 1) / TBM, nt = (p.N + BN - 1) / BN;
  // Tile schedule.  CL == 1: CTA-granular, tile t = blockIdx.x + i * gridDim.x, m fastest.
  // CL == 4: a 2x2 cluster walks 2x2 super-tiles; rank r -> (ci = r >> 1 along M, cj = r & 1 along N).
  const uint32_t crank = (CL == 1) ? 0u : cluster_ctarank();
  const int ci = (int)(crank >> 1), cj = (int)(crank & 1);
  const int smt = (CL == 1) ? mt : (mt + 1) / 2, snt = (CL == 1) ? nt : (nt + 1) / 2;
  const long long ntiles = (long long)smt * snt * p.batch;
  const long long t_first = (CL == 1) ? (long long)blockIdx.x : (long long)cluster_id_x();
  const long long t_step = (CL == 1) ? (long long)gridDim.x : (long long)num_clusters_x();
  auto tile_coords = [&](long long t, int& m0, int& n0, int& z) {
    const int tm = (int)(t % smt), tn = (int)((t / smt) % snt);
    z = (int)(t / ((long long)smt * snt));
    m0 = ((CL == 1) ? tm : 2 * tm + ci) * TBM;
    n0 = ((CL == 1) ? tn : 2 * tn + cj) * BN;
  };
  // multicast masks: A tile is shared by the two CTAs with the same ci, B tile by the two with the same cj
  const uint16_t a_mask = (uint16_t)(0x3u << (2 * ci)), b_mask = (uint16_t)((1u << cj) | (1u << (cj + 2)));
  const uint16_t e_mask = (uint16_t)((1u << crank) | (1u << (crank ^ 1u)) | (1u << (crank ^ 2u)));

  if (warp == 0 && lane == 0) {
    // a stage is written by this CTA and (CL == 4) by its A- and B-partners: all three consumers must release it
    for (int s = 0; s < SL::STAGES; ++s) { mbar_init(full_bar(s), 1); mbar_init(empty_bar(s), CL == 1 ? 1 : 3); }
    for (int b = 0; b < 2; ++b) { mbar_init(tfull_bar(b), 1); mbar_init(tempty_bar(b), 8); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  if (warp == 2) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot),
                 "r"((uint32_t)TMEM_COLS));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::);
  }
  tc_fence_before();
  __syncthreads();
  if (CL != 1) cluster_sync_all();          // partner barriers are initialised before any multicast / remote arrive
  tc_fence_after();
  uint32_t tmem_base;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot));

  if (warp == 0) {
    // ================= TMA producer (warp-uniform loop, one elected lane issues) =================
    {
      uint32_t it = 0;   // k-block counter across all tiles of this CTA
      for (long long t = t_first; t < ntiles; t += t_step) {
        int m0, n0, z;
        tile_coords(t, m0, n0, z);
        for (int kb = 0; kb < nk; ++kb, ++it) {
          const int s = it % SL::STAGES;
          const uint32_t ph = (it / SL::STAGES) & 1;
          mbar_wait(empty_bar(s), ph ^ 1);
          const uint32_t st = smem_base + s * SL::STAGE;
          if (elect_one()) {
            if (p.dbg & 4) {
              mbar_arrive(full_bar(s));
            } else {
              const bool second = kb >= nk1;
              const bool no_blo = skip_b1lo && !second;
              mbar_arrive_expect_tx(full_bar(s), SL::STAGE - (no_blo ? SL::B_TILE : 0));
              const int k0 = (second ? kb - nk1 : kb) * TBK;
              const CUtensorMap* ah = second ? &mA2h : &mA1h;
              const CUtensorMap* al = second ? &mA2l : &mA1l;
              const CUtensorMap* bh = second ? &mB2h : &mB1h;
              const CUtensorMap* bl = second ? &mB2l : &mB1l;
              const int za = (second ? p.a2_batched : p.a1_batched) ? z : 0;
              const int zb = (second ? p.b2_batched : p.b1_batched) ? z : 0;
              load_operand<A_K, TBM, CL>(st, ah, full_bar(s), k0, m0, za, cj, a_mask);
              load_operand<A_K, TBM, CL>(st + SL::A_TILE, al, full_bar(s), k0, m0, za, cj, a_mask);
              load_operand<B_K, BN, CL>(st + 2 * SL::A_TILE, bh, full_bar(s), k0, n0, zb, ci, b_mask);
              if (!no_blo) load_operand<B_K, BN, CL>(st + 2 * SL::A_TILE + SL::B_TILE, bl, full_bar(s), k0, n0, zb, ci, b_mask);
            }
          }
          __syncwarp();
        }
      }
    }
  } else if (warp == 1) {
    // ================= MMA issuer (warp-uniform loop, one elected lane issues) =================
    {
      constexpr uint32_t idesc = make_idesc(TBM, BN, !A_K, !B_K);
      // K-major SW128: 8-row groups 1024 B apart (SBO); k sub-step (8 tf32) = +32 B inside the 128 B row.
      // MN-major SW128_BASE32B: 32-element MN chunks TBK*128 B apart (LBO), 4-row k groups 512 B apart (SBO);
      //   k sub-step (8 rows) = +1024 B.
      constexpr uint32_t A_LBO = A_K ? 16 : TBK * 128, B_LBO = B_K ? 16 : TBK * 128;
      constexpr uint32_t A_SBO = A_K ? 1024 : 512, B_SBO = B_K ? 1024 : 512;
      constexpr uint32_t A_LT = A_K ? 2 : 1, B_LT = B_K ? 2 : 1;
      constexpr uint32_t A_KSTEP = A_K ? 32 : 1024, B_KSTEP = B_K ? 32 : 1024;
      uint32_t it = 0, ck = 0;   // k-block / chunk counters across all tiles
      for (long long t = t_first; t < ntiles; t += t_step) {
        for (int c = 0; c < nchunks; ++c, ++ck) {
          const uint32_t buf = ck & 1, cph = (ck >> 1) & 1;
          mbar_wait(tempty_bar(buf), cph ^ 1);           // drain warps have emptied this TMEM buffer
          tc_fence_after();
          const uint32_t t_small = tmem_base + buf * (2 * BN), t_main = t_small + BN;
          const int kb_end = (c + 1) * KCr < nk ? (c + 1) * KCr : nk;
          for (int kb = c * KCr; kb < kb_end; ++kb, ++it) {
            const int s = it % SL::STAGES;
            const uint32_t ph = (it / SL::STAGES) & 1;
            mbar_wait(full_bar(s), ph);
            tc_fence_after();
            const uint32_t st = smem_base + s * SL::STAGE;
            const uint32_t a_hi = st, a_lo = st + SL::A_TILE, b_hi = st + 2 * SL::A_TILE, b_lo = b_hi + SL::B_TILE;
            if (elect_one()) {
#pragma unroll
              for (int j = 0; j < TBK / UMMA_K; ++j) {
                const uint64_t dah = make_smem_desc(a_hi + j * A_KSTEP, A_LBO, A_SBO, A_LT);
                const uint64_t dal = make_smem_desc(a_lo + j * A_KSTEP, A_LBO, A_SBO, A_LT);
                const uint64_t dbh = make_smem_desc(b_hi + j * B_KSTEP, B_LBO, B_SBO, B_LT);
                const uint64_t dbl = make_smem_desc(b_lo + j * B_KSTEP, B_LBO, B_SBO, B_LT);
                const uint32_t acc = (kb != c * KCr || j != 0) ? 1u : 0u;   // first MMA of a chunk overwrites
                const bool no_blo = skip_b1lo && kb < nk1;
                if (p.merge) {
                  umma_tf32(t_main, dal, dbh, idesc, acc);
                  if (!no_blo) umma_tf32(t_main, dah, dbl, idesc, 1);
                  umma_tf32(t_main, dah, dbh, idesc, 1);
                } else {
                  if (!(p.dbg & 2)) {
                    umma_tf32(t_small, dal, dbh, idesc, acc);
                    if (!no_blo) umma_tf32(t_small, dah, dbl, idesc, 1);
                  }
                  umma_tf32(t_main, dah, dbh, idesc, acc);
                }
              }
              // frees the smem stage when these MMAs retire (for every CTA that writes into it)
              if (CL == 1) umma_commit(empty_bar(s)); else umma_commit_mc(empty_bar(s), e_mask);
            }
            __syncwarp();
          }
          if (elect_one()) umma_commit(tfull_bar(buf));            // chunk complete
          __syncwarp();
        }
      }
    }
  } else if (warp >= 4) {
    // ================= drain + epilogue warps (8): warp -> TMEM lane quarter q, column half h =================
    const int q = warp & 3;                   // TMEM lane quarter this warp may access (warp id % 4)
    const int h = (warp - 4) >> 2;            // which 64-column half of the tile this warp owns
    constexpr int HC = BN / 2;                // 64 columns per warp
    float* stg = reinterpret_cast<float*>(smem_raw + (stg_base - smem_u32(smem_raw))) + (warp - 4) * 32 * STG_LD;
    uint32_t ck = 0;
    for (long long t = t_first; t < ntiles; t += t_step) {
      int m0, n0, z;
      tile_coords(t, m0, n0, z);
      float acc[HC];
#pragma unroll
      for (int i = 0; i < HC; ++i) acc[i] = 0.f;
      for (int c = 0; c < nchunks; ++c, ++ck) {
        const uint32_t buf = ck & 1, cph = (ck >> 1) & 1;
        mbar_wait(tfull_bar(buf), cph);
        tc_fence_after();
        const uint32_t tl = tmem_base + ((uint32_t)(q * 32) << 16) + buf * (2 * BN) + h * HC;
#pragma unroll
        for (int cc = 0; cc < HC / 32; ++cc) {
          float w[32];
          if (!p.merge) {
            tmem_ld32(tl + (uint32_t)(cc * 32), w);              // cross terms
#pragma unroll
            for (int i = 0; i < 32; ++i) acc[cc * 32 + i] += w[i];
          }
          tmem_ld32(tl + (uint32_t)(BN + cc * 32), w);         // main terms
#pragma unroll
          for (int i = 0; i < 32; ++i) acc[cc * 32 + i] += w[i];
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(tempty_bar(buf));           // buffer may be overwritten by the next chunk
      }
      // ---- tile epilogue: registers -> smem (own 32x32 blocks) -> coalesced fused epilogue + store ----
      tile_epilogue<2, 8>(p, acc, stg, m0, n0, z, q, h, lane);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (CL != 1) cluster_sync_all();          // no multicast write / remote arrive may target a CTA that has exited
  if (warp == 2) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)TMEM_COLS));
  }
}

// ============================================================================================================
// 2-CTA variant (cta_group::2): a CTA pair computes a 256 x BN2 tile.  Each CTA stages its own 128 A rows and HALF
// of the B tile (BN2/2 rows); the tensor core of the pair reads B halves from both shared memories, so per-CTA
// shared-memory operand traffic drops from 8 KB to 6 KB per MMA and TMA fill from 64 KB to 48 KB per k-block
// (the 1-CTA kernel saturates the 128 B/clk shared-memory port at ~60 % tensor utilisation).
// MMAs are issued by the leader CTA (rank 0) only; TMA loads of both CTAs credit the leader's full barrier; the
// leader's commits are multicast to both CTAs' empty / tmem-full barriers; the peer's drain warps release the
// TMEM buffers on the leader's tmem-empty barrier.  Everything else (chunked drain, epilogue) is per-CTA as above.
// ============================================================================================================
template <int BN2>
struct SmemLayout2 {
  static constexpr int A_TILE = TBM * TBK * 4;          // 16 KB
  static constexpr int B_TILE = (BN2 / 2) * TBK * 4;    // half of the B tile
  static constexpr int STAGE = 2 * A_TILE + 2 * B_TILE;
  static constexpr int STAGES = 4;
  static constexpr int STAGING = 8 * 32 * STG_LD * 4;
  static constexpr int BYTES = STAGES * STAGE + STAGING + 1024 + 256;
};

template <bool KMAJOR, int ROWS>
__device__ __forceinline__ void load_operand_2sm(uint32_t dst, const CUtensorMap* map, uint32_t bar, int k0, int row0, int z) {
  if (KMAJOR) {
    tma_load_3d_2sm(dst, map, bar, k0, row0, z);
  } else {
#pragma unroll
    for (int c = 0; c < ROWS / 32; ++c) tma_load_3d_2sm(dst + c * (TBK * 128), map, bar, row0 + 32 * c, k0, z);
  }
}

template <int BN2, bool A_K, bool B_K>
__global__ void __launch_bounds__(TC_THREADS, 1)
gemm_tc2_kernel(const __grid_constant__ CUtensorMap mA1h, const __grid_constant__ CUtensorMap mA1l,
                const __grid_constant__ CUtensorMap mB1h, const __grid_constant__ CUtensorMap mB1l,
                const __grid_constant__ CUtensorMap mA2h, const __grid_constant__ CUtensorMap mA2l,
                const __grid_constant__ CUtensorMap mB2h, const __grid_constant__ CUtensorMap mB2l, TcParams p) {
  using SL = SmemLayout2<BN2>;
  static_assert(BN2 == 128, "drain register accumulators are sized for 128 output columns per CTA");
  constexpr int BN = BN2;                 // output columns per CTA
  constexpr int TMEM_COLS = 512;          // 2 buffers x (cross-term tile + main tile) x BN columns
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t stg_base = smem_base + SL::STAGES * SL::STAGE;
  const uint32_t bar_base = stg_base + SL::STAGING;
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (SL::STAGES + s); };
  auto tfull_bar = [&](int b) { return bar_base + 8u * (2 * SL::STAGES + b); };
  auto tempty_bar = [&](int b) { return bar_base + 8u * (2 * SL::STAGES + 2 + b); };
  const uint32_t tmem_slot = bar_base + 8u * (2 * SL::STAGES + 4);

  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);   // provably warp-uniform
  const int lane = threadIdx.x & 31;
  const uint32_t crank = cluster_ctarank();          // 0 = leader
  const int nk1 = (p.K1 + TBK - 1) / TBK, nk2 = (p.K2 + TBK - 1) / TBK;
  const int nk = nk1 + nk2;
  const bool skip_b1lo = p.b1_lo_nz != nullptr && *reinterpret_cast<const volatile int*>(p.b1_lo_nz) == 0;
  const int KCr = p.kc;
  const int nchunks = (nk + KCr - 1) / KCr;
  const int mpt = (p.M + 2 * TBM - 1) / (2 * TBM), nt = (p.N + BN - 1) / BN;
  const long long ntiles = (long long)mpt * nt * p.batch;
  const long long t_first = (long long)cluster_id_x(), t_step = (long long)num_clusters_x();
  auto tile_coords = [&](long long t, int& m0, int& n0, int& z) {
    z = (int)(t / ((long long)mpt * nt));
    m0 = ((int)(t % mpt) * 2 + (int)crank) * TBM;
    n0 = (int)((t / mpt) % nt) * BN;
  };

  if (warp == 0 && lane == 0) {
    for (int s = 0; s < SL::STAGES; ++s) { mbar_init(full_bar(s), 1); mbar_init(empty_bar(s), 1); }
    // tmem-full: one multicast commit from the leader; tmem-empty (used on the leader): 8 drain warps x 2 CTAs
    for (int b = 0; b < 2; ++b) { mbar_init(tfull_bar(b), 1); mbar_init(tempty_bar(b), 16); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  if (warp == 2) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot),
                 "r"((uint32_t)TMEM_COLS));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::);
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  tc_fence_after();
  uint32_t tmem_base;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot));

  if (warp == 0) {
    // ================= TMA producer (both CTAs; warp-uniform loop, one elected lane issues) =================
    {
      uint32_t it = 0;
      for (long long t = t_first; t < ntiles; t += t_step) {
        int m0, n0, z;
        tile_coords(t, m0, n0, z);
        const int nb0 = n0 + (int)crank * (BN / 2);       // this CTA stages B rows [nb0, nb0 + BN/2)
        for (int kb = 0; kb < nk; ++kb, ++it) {
          const int s = it % SL::STAGES;
          const uint32_t ph = (it / SL::STAGES) & 1;
          mbar_wait(empty_bar(s), ph ^ 1);
          const uint32_t st = smem_base + s * SL::STAGE;
          if (elect_one()) {
            if (p.dbg & 4) {
              if (crank == 0) mbar_arrive(full_bar(s));
            } else {
              const bool second = kb >= nk1;
              const bool no_blo = skip_b1lo && !second;
              if (crank == 0) mbar_arrive_expect_tx(full_bar(s), 2 * (SL::STAGE - (no_blo ? SL::B_TILE : 0)));   // bytes of BOTH CTAs
              const int k0 = (second ? kb - nk1 : kb) * TBK;
              const CUtensorMap* ah = second ? &mA2h : &mA1h;
              const CUtensorMap* al = second ? &mA2l : &mA1l;
              const CUtensorMap* bh = second ? &mB2h : &mB1h;
              const CUtensorMap* bl = second ? &mB2l : &mB1l;
              const int za = (second ? p.a2_batched : p.a1_batched) ? z : 0;
              const int zb = (second ? p.b2_batched : p.b1_batched) ? z : 0;
              load_operand_2sm<A_K, TBM>(st, ah, full_bar(s), k0, m0, za);
              load_operand_2sm<A_K, TBM>(st + SL::A_TILE, al, full_bar(s), k0, m0, za);
              load_operand_2sm<B_K, BN / 2>(st + 2 * SL::A_TILE, bh, full_bar(s), k0, nb0, zb);
              if (!no_blo) load_operand_2sm<B_K, BN / 2>(st + 2 * SL::A_TILE + SL::B_TILE, bl, full_bar(s), k0, nb0, zb);
            }
          }
          __syncwarp();
        }
      }
    }
  } else if (warp == 1) {
    // ================= MMA issuer: LEADER CTA only (warp-uniform loop, one elected lane issues) =================
    if (crank == 0) {
      constexpr uint32_t idesc = make_idesc(2 * TBM, BN, !A_K, !B_K);
      constexpr uint32_t A_LBO = A_K ? 16 : TBK * 128, B_LBO = B_K ? 16 : TBK * 128;
      constexpr uint32_t A_SBO = A_K ? 1024 : 512, B_SBO = B_K ? 1024 : 512;
      constexpr uint32_t A_LT = A_K ? 2 : 1, B_LT = B_K ? 2 : 1;
      constexpr uint32_t A_KSTEP = A_K ? 32 : 1024, B_KSTEP = B_K ? 32 : 1024;
      uint32_t it = 0, ck = 0;
      for (long long t = t_first; t < ntiles; t += t_step) {
        for (int c = 0; c < nchunks; ++c, ++ck) {
          const uint32_t buf = ck & 1, cph = (ck >> 1) & 1;
          mbar_wait(tempty_bar(buf), cph ^ 1);           // both CTAs' drain warps have emptied this TMEM buffer
          tc_fence_after();
          const uint32_t t_small = tmem_base + buf * (2 * BN), t_main = t_small + BN;
          const int kb_end = (c + 1) * KCr < nk ? (c + 1) * KCr : nk;
          for (int kb = c * KCr; kb < kb_end; ++kb, ++it) {
            const int s = it % SL::STAGES;
            const uint32_t ph = (it / SL::STAGES) & 1;
            mbar_wait(full_bar(s), ph);
            tc_fence_after();
            const uint32_t st = smem_base + s * SL::STAGE;
            const uint32_t a_hi = st, a_lo = st + SL::A_TILE, b_hi = st + 2 * SL::A_TILE, b_lo = b_hi + SL::B_TILE;
            if (elect_one()) {
#pragma unroll
              for (int j = 0; j < TBK / UMMA_K; ++j) {
                const uint64_t dah = make_smem_desc(a_hi + j * A_KSTEP, A_LBO, A_SBO, A_LT);
                const uint64_t dal = make_smem_desc(a_lo + j * A_KSTEP, A_LBO, A_SBO, A_LT);
                const uint64_t dbh = make_smem_desc(b_hi + j * B_KSTEP, B_LBO, B_SBO, B_LT);
                const uint64_t dbl = make_smem_desc(b_lo + j * B_KSTEP, B_LBO, B_SBO, B_LT);
                const uint32_t acc = (kb != c * KCr || j != 0) ? 1u : 0u;
                const bool no_blo = skip_b1lo && kb < nk1;
                if (p.merge) {
                  umma_tf32_2sm(t_main, dal, dbh, idesc, acc);
                  if (!no_blo) umma_tf32_2sm(t_main, dah, dbl, idesc, 1);
                  umma_tf32_2sm(t_main, dah, dbh, idesc, 1);
                } else {
                  if (!(p.dbg & 2)) {
                    umma_tf32_2sm(t_small, dal, dbh, idesc, acc);
                    if (!no_blo) umma_tf32_2sm(t_small, dah, dbl, idesc, 1);
                  }
                  umma_tf32_2sm(t_main, dah, dbh, idesc, acc);
                }
              }
              umma_commit_2sm(empty_bar(s), 0x3);     // stage free in both CTAs
            }
            __syncwarp();
          }
          if (elect_one()) umma_commit_2sm(tfull_bar(buf), 0x3);     // chunk complete: wake both CTAs' drain warps
          __syncwarp();
        }
      }
    }
  } else if (warp >= 4) {
    // ================= drain + epilogue warps (per CTA: its own 128 rows x BN columns) =================
    const int q = warp & 3;
    const int h = (warp - 4) >> 2;
    constexpr int HC = BN / 2;
    float* stg = reinterpret_cast<float*>(smem_raw + (stg_base - smem_u32(smem_raw))) + (warp - 4) * 32 * STG_LD;
    uint32_t ck = 0;
    for (long long t = t_first; t < ntiles; t += t_step) {
      int m0, n0, z;
      tile_coords(t, m0, n0, z);
      float acc[HC];
#pragma unroll
      for (int i = 0; i < HC; ++i) acc[i] = 0.f;
      for (int c = 0; c < nchunks; ++c, ++ck) {
        const uint32_t buf = ck & 1, cph = (ck >> 1) & 1;
        mbar_wait(tfull_bar(buf), cph);
        tc_fence_after();
        const uint32_t tl = tmem_base + ((uint32_t)(q * 32) << 16) + buf * (2 * BN) + h * HC;
#pragma unroll
        for (int cc = 0; cc < HC / 32; ++cc) {
          float w[32];
          if (!p.merge) {
            tmem_ld32(tl + (uint32_t)(cc * 32), w);
#pragma unroll
            for (int i = 0; i < 32; ++i) acc[cc * 32 + i] += w[i];
          }
          tmem_ld32(tl + (uint32_t)(BN + cc * 32), w);
#pragma unroll
          for (int i = 0; i < 32; ++i) acc[cc * 32 + i] += w[i];
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) {
          if (crank == 0) mbar_arrive(tempty_bar(buf)); else mbar_arrive_remote(tempty_bar(buf), 0);
        }
      }
      tile_epilogue<2, 8>(p, acc, stg, m0, n0, z, q, h, lane);
    }
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  if (warp == 2) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)TMEM_COLS));
  }
}

// ============================================================================================================
// Wide CTA-pair variant: a pair computes a 256 x 256 tile (UMMA 256 x 256 x 8, cta_group::2).  Per CTA and MMA the
// tensor core reads 4 KB of A and 4 KB of B from shared memory for 128 clk of work: half the operand traffic of the
// 128-wide tiles, which are operand-fetch bound (tools/mma_probe.cu: 139 clk per MMA here = 3765 FLOP/clk/SM, the nominal
// TF32 rate, against 3021 for 128 x 128).  The 512 TMEM columns hold ONE cross-term tile and ONE main tile (256 columns
// each), so there is no TMEM double buffering: the cross terms accumulate over the whole K loop (their truncation error
// is 2^-11 down), the main terms are folded into fp32 registers every KC k-blocks and the first k-block of the next chunk
// issues its cross-term MMAs while that drain is in flight.  Drain warps hold 32 x 128 accumulators each, so the kernel
// re-balances registers between the control warpgroup and the two drain warpgroups with setmaxnreg.
// ============================================================================================================
constexpr int WN = 256;                 // pair-tile columns (per CTA: 128 rows x 256 columns of output)
constexpr int W_REG_CTRL = 40, W_REG_DRAIN = 224;

struct SmemLayoutW {
  static constexpr int A_TILE = TBM * TBK * 4;          // 16 KB
  static constexpr int B_TILE = (WN / 2) * TBK * 4;     // half of the 256-column B tile: 16 KB
  static constexpr int STAGE = 2 * A_TILE + 2 * B_TILE; // 64 KB
  static constexpr int STAGES = 3;
  static constexpr int STAGING = 8 * 32 * STG_LD * 4;
  static constexpr int BYTES = STAGES * STAGE + STAGING + 1024 + 256;
};

template <bool A_K, bool B_K>
__global__ void __launch_bounds__(TC_THREADS, 1)
gemm_tc2w_kernel(const __grid_constant__ CUtensorMap mA1h, const __grid_constant__ CUtensorMap mA1l,
                 const __grid_constant__ CUtensorMap mB1h, const __grid_constant__ CUtensorMap mB1l,
                 const __grid_constant__ CUtensorMap mA2h, const __grid_constant__ CUtensorMap mA2l,
                 const __grid_constant__ CUtensorMap mB2h, const __grid_constant__ CUtensorMap mB2l, TcParams p) {
  using SL = SmemLayoutW;
  constexpr int TMEM_COLS = 512;          // [0, 256): cross terms (whole tile), [256, 512): main terms (one chunk)
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t stg_base = smem_base + SL::STAGES * SL::STAGE;
  const uint32_t bar_base = stg_base + SL::STAGING;
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (SL::STAGES + s); };
  const uint32_t mfull_bar = bar_base + 8u * (2 * SL::STAGES + 0);    // main tile of a chunk complete
  const uint32_t mempty_bar = bar_base + 8u * (2 * SL::STAGES + 1);   // main tile drained (leader's barrier)
  const uint32_t cfull_bar = bar_base + 8u * (2 * SL::STAGES + 2);    // cross tile of a tile complete
  const uint32_t cempty_bar = bar_base + 8u * (2 * SL::STAGES + 3);   // cross tile drained (leader's barrier)
  const uint32_t tmem_slot = bar_base + 8u * (2 * SL::STAGES + 4);

  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);
  const int lane = threadIdx.x & 31;
  const uint32_t crank = cluster_ctarank();
  const int nk1 = (p.K1 + TBK - 1) / TBK, nk2 = (p.K2 + TBK - 1) / TBK;
  const int nk = nk1 + nk2;
  const bool skip_b1lo = p.b1_lo_nz != nullptr && *reinterpret_cast<const volatile int*>(p.b1_lo_nz) == 0;
  const int KCr = p.kc;
  const int nchunks = (nk + KCr - 1) / KCr;
  const int mpt = (p.M + 2 * TBM - 1) / (2 * TBM), nt = (p.N + WN - 1) / WN;
  const long long ntiles = (long long)mpt * nt * p.batch;
  const long long t_first = (long long)cluster_id_x(), t_step = (long long)num_clusters_x();
  auto tile_coords = [&](long long t, int& m0, int& n0, int& z) {
    z = (int)(t / ((long long)mpt * nt));
    m0 = ((int)(t % mpt) * 2 + (int)crank) * TBM;
    n0 = (int)((t / mpt) % nt) * WN;
  };

  if (warp == 0 && lane == 0) {
    for (int s = 0; s < SL::STAGES; ++s) { mbar_init(full_bar(s), 1); mbar_init(empty_bar(s), 1); }
    mbar_init(mfull_bar, 1); mbar_init(cfull_bar, 1);
    mbar_init(mempty_bar, 16); mbar_init(cempty_bar, 16);       // 8 drain warps x 2 CTAs
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  if (warp == 2) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot),
                 "r"((uint32_t)TMEM_COLS));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::);
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  tc_fence_after();
  uint32_t tmem_base;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot));

  if (warp < 4) {
    asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(W_REG_CTRL));
    if (warp == 0) {
      // ================= TMA producer (both CTAs) =================
      uint32_t it = 0;
      for (long long t = t_first; t < ntiles; t += t_step) {
        int m0, n0, z;
        tile_coords(t, m0, n0, z);
        const int nb0 = n0 + (int)crank * (WN / 2);       // this CTA stages B rows [nb0, nb0 + 128)
        for (int kb = 0; kb < nk; ++kb, ++it) {
          const int s = it % SL::STAGES;
          const uint32_t ph = (it / SL::STAGES) & 1;
          mbar_wait(empty_bar(s), ph ^ 1);
          const uint32_t st = smem_base + s * SL::STAGE;
          if (elect_one()) {
            if (p.dbg & 4) {
              if (crank == 0) mbar_arrive(full_bar(s));
            } else {
              const bool second = kb >= nk1;
              const bool no_blo = skip_b1lo && !second;
              if (crank == 0) mbar_arrive_expect_tx(full_bar(s), 2 * (SL::STAGE - (no_blo ? SL::B_TILE : 0)));
              const int k0 = (second ? kb - nk1 : kb) * TBK;
              const CUtensorMap* ah = second ? &mA2h : &mA1h;
              const CUtensorMap* al = second ? &mA2l : &mA1l;
              const CUtensorMap* bh = second ? &mB2h : &mB1h;
              const CUtensorMap* bl = second ? &mB2l : &mB1l;
              const int za = (second ? p.a2_batched : p.a1_batched) ? z : 0;
              const int zb = (second ? p.b2_batched : p.b1_batched) ? z : 0;
              load_operand_2sm<A_K, TBM>(st, ah, full_bar(s), k0, m0, za);
              load_operand_2sm<A_K, TBM>(st + SL::A_TILE, al, full_bar(s), k0, m0, za);
              load_operand_2sm<B_K, WN / 2>(st + 2 * SL::A_TILE, bh, full_bar(s), k0, nb0, zb);
              if (!no_blo) load_operand_2sm<B_K, WN / 2>(st + 2 * SL::A_TILE + SL::B_TILE, bl, full_bar(s), k0, nb0, zb);
            }
          }
          __syncwarp();
        }
      }
    } else if (warp == 1 && crank == 0) {
      // ================= MMA issuer: leader CTA only =================
      constexpr uint32_t idesc = make_idesc(2 * TBM, WN, !A_K, !B_K);
      constexpr uint32_t A_LBO = A_K ? 16 : TBK * 128, B_LBO = B_K ? 16 : TBK * 128;
      constexpr uint32_t A_SBO = A_K ? 1024 : 512, B_SBO = B_K ? 1024 : 512;
      constexpr uint32_t A_LT = A_K ? 2 : 1, B_LT = B_K ? 2 : 1;
      constexpr uint32_t A_KSTEP = A_K ? 32 : 1024, B_KSTEP = B_K ? 32 : 1024;
      const uint32_t t_cross = tmem_base, t_main = tmem_base + WN;
      uint32_t it = 0, ck = 0, tl = 0;     // k-block / chunk / tile counters
      for (long long t = t_first; t < ntiles; t += t_step, ++tl) {
        for (int c = 0; c < nchunks; ++c, ++ck) {
          const int kb_end = (c + 1) * KCr < nk ? (c + 1) * KCr : nk;
          for (int kb = c * KCr; kb < kb_end; ++kb, ++it) {
            const int s = it % SL::STAGES;
            const uint32_t ph = (it / SL::STAGES) & 1;
            mbar_wait(full_bar(s), ph);
            if (kb == 0) mbar_wait(cempty_bar, (tl & 1) ^ 1);        // previous tile's cross terms have been read out
            tc_fence_after();
            const uint32_t st = smem_base + s * SL::STAGE;
            const uint32_t a_hi = st, a_lo = st + SL::A_TILE, b_hi = st + 2 * SL::A_TILE, b_lo = b_hi + SL::B_TILE;
            if (elect_one()) {
#pragma unroll
              for (int j = 0; j < TBK / UMMA_K; ++j) {      // cross terms first: they do not need the main tile
                const uint64_t dah = make_smem_desc(a_hi + j * A_KSTEP, A_LBO, A_SBO, A_LT);
                const uint64_t dal = make_smem_desc(a_lo + j * A_KSTEP, A_LBO, A_SBO, A_LT);
                const uint64_t dbh = make_smem_desc(b_hi + j * B_KSTEP, B_LBO, B_SBO, B_LT);
                const uint64_t dbl = make_smem_desc(b_lo + j * B_KSTEP, B_LBO, B_SBO, B_LT);
                if (!(p.dbg & 2)) {
                  umma_tf32_2sm(t_cross, dal, dbh, idesc, (kb != 0 || j != 0) ? 1u : 0u);
                  if (!(skip_b1lo && kb < nk1)) umma_tf32_2sm(t_cross, dah, dbl, idesc, 1);
                }
              }
            }
            __syncwarp();
            if (kb == c * KCr) {                                       // the previous chunk's main tile has been folded
              mbar_wait(mempty_bar, (ck & 1) ^ 1);
              tc_fence_after();
            }
            if (elect_one()) {
#pragma unroll
              for (int j = 0; j < TBK / UMMA_K; ++j) {
                const uint64_t dah = make_smem_desc(a_hi + j * A_KSTEP, A_LBO, A_SBO, A_LT);
                const uint64_t dbh = make_smem_desc(b_hi + j * B_KSTEP, B_LBO, B_SBO, B_LT);
                umma_tf32_2sm(t_main, dah, dbh, idesc, (kb != c * KCr || j != 0) ? 1u : 0u);
              }
              umma_commit_2sm(empty_bar(s), 0x3);
            }
            __syncwarp();
          }
          if (elect_one()) {
            umma_commit_2sm(mfull_bar, 0x3);
            if (c == nchunks - 1) umma_commit_2sm(cfull_bar, 0x3);
          }
          __syncwarp();
        }
      }
    }
  } else {
    asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(W_REG_DRAIN));
    // ================= drain + epilogue warps (per CTA: its own 128 rows x 256 columns) =================
    const int q = warp & 3;
    const int h = (warp - 4) >> 2;            // which 128-column half
    constexpr int HC = WN / 2;                // 128 columns per warp
    float* stg = reinterpret_cast<float*>(smem_raw + (stg_base - smem_u32(smem_raw))) + (warp - 4) * 32 * STG_LD;
    uint32_t ck = 0, tl = 0;
    for (long long t = t_first; t < ntiles; t += t_step, ++tl) {
      int m0, n0, z;
      tile_coords(t, m0, n0, z);
      float acc[HC];
#pragma unroll
      for (int i = 0; i < HC; ++i) acc[i] = 0.f;
      const uint32_t tbase = tmem_base + ((uint32_t)(q * 32) << 16) + h * HC;
      for (int c = 0; c < nchunks; ++c, ++ck) {
        mbar_wait(mfull_bar, ck & 1);
        tc_fence_after();
#pragma unroll
        for (int cc = 0; cc < HC / 32; ++cc) {
          float w[32];
          tmem_ld32(tbase + (uint32_t)(WN + cc * 32), w);         // main terms of this chunk
#pragma unroll
          for (int i = 0; i < 32; ++i) acc[cc * 32 + i] += w[i];
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) {
          if (crank == 0) mbar_arrive(mempty_bar); else mbar_arrive_remote(mempty_bar, 0);
        }
      }
      mbar_wait(cfull_bar, tl & 1);
      tc_fence_after();
      if (!(p.dbg & 2)) {
#pragma unroll
        for (int cc = 0; cc < HC / 32; ++cc) {
          float w[32];
          tmem_ld32(tbase + (uint32_t)(cc * 32), w);               // cross terms of the whole tile
#pragma unroll
          for (int i = 0; i < 32; ++i) acc[cc * 32 + i] += w[i];
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) {
        if (crank == 0) mbar_arrive(cempty_bar); else mbar_arrive_remote(cempty_bar, 0);
      }
      tile_epilogue<HC / 32, 32>(p, acc, stg, m0, n0, z, q, h, lane);
    }
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  if (warp == 2) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)TMEM_COLS));
  }
}

// ---- TF32 hi/lo split --------------------------------------------------------------------------------------
// src element (z, r, c) at src + z*sz_src + r*ld_src + c  ->  hi/lo (z, r, c) at z*sz_dst + r*ld_dst + c, c < ld_dst
// (columns in [cols, ld_dst) are zero-filled)
__global__ void tf32_split_kernel(const float* __restrict__ src, long long sz_src, long long ld_src,
                                  float* __restrict__ hi, float* __restrict__ lo, long long sz_dst, long long ld_dst,
                                  long long cols, int* __restrict__ nz) {
  int any = 0;
  const long long r = blockIdx.y, z = blockIdx.z;
  const float* s = src + z * sz_src + r * ld_src;
  float* h = hi + z * sz_dst + r * ld_dst;
  float* l = lo + z * sz_dst + r * ld_dst;
  for (long long c = blockIdx.x * (long long)blockDim.x + threadIdx.x; c < ld_dst; c += (long long)gridDim.x * blockDim.x) {
    float x = c < cols ? s[c] : 0.f;
    float xh = tf32_rna(x);
    const float xl = tf32_rna(x - xh);
    h[c] = xh;
    l[c] = xl;
    any |= (xl != 0.f);
  }
  if (nz && __syncthreads_or(any) && threadIdx.x == 0) atomicOr(nz, 1);
}

// Contiguous blocks (ld_src == ld_dst == cols, the probe blocks V[b, woff : woff + in*out]): one flat grid-stride pass
// per batch entry, 4 independent coalesced loads in flight per thread (the source is only 4-byte aligned: b*D + woff).
__global__ void __launch_bounds__(256) tf32_split_flat_kernel(const float* __restrict__ src, long long sz_src,
                                                              float* __restrict__ hi, float* __restrict__ lo,
                                                              long long sz_dst, long long n, int* __restrict__ nz) {
  int any = 0;
  const long long z = blockIdx.y;
  const float* s = src + z * sz_src;
  float* h = hi + z * sz_dst;
  float* l = lo + z * sz_dst;
  const long long stride = (long long)gridDim.x * 256;
  long long i = blockIdx.x * 256ll + threadIdx.x;
  for (; i + 3 * stride < n; i += 4 * stride) {
    float x[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) x[u] = __ldg(s + i + u * stride);
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const float xh = tf32_rna(x[u]);
      const float xl = tf32_rna(x[u] - xh);
      h[i + u * stride] = xh;
      l[i + u * stride] = xl;
      any |= (xl != 0.f);
    }
  }
  for (; i < n; i += stride) {
    const float x = __ldg(s + i), xh = tf32_rna(x);
    const float xl = tf32_rna(x - xh);
    h[i] = xh;
    l[i] = xl;
    any |= (xl != 0.f);
  }
  if (nz && __syncthreads_or(any) && threadIdx.x == 0) atomicOr(nz, 1);
}

// ---- host: tensor maps ---------------------------------------------------------------------------------------
EncodeTiledFn get_encode() {
  static EncodeTiledFn fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = (EncodeTiledFn)p;
  });
  return fn;
}

// kmajor: element (row r, k) at ptr + r*ld + k;  box {32 k, box_rows rows, 1}
// mn-major: element (k, col c) at ptr + k*ld + c; box {32 cols, 32 k, 1} (one box per 32-column chunk)
int make_map(CUtensorMap* map, const float* ptr, bool kmajor, int64_t rows_or_cols, int64_t K, int64_t ld, int64_t sz,
             int64_t batch, int box_rows) {
  EncodeTiledFn enc = get_encode();
  if (!enc) { set_error("cuTensorMapEncodeTiled unavailable"); return LIP_ERR_UNSUPPORTED; }
  LIP_REQUIRE(((uintptr_t)ptr & 15) == 0 && ld % 4 == 0 && sz % 4 == 0, "gemm_tc: operand not 16-byte aligned (ld=%lld sz=%lld)",
              (long long)ld, (long long)sz);
  CUresult r;
  const cuuint64_t bstride = (cuuint64_t)(batch > 1 ? sz : (kmajor ? rows_or_cols * ld : K * ld)) * 4;
  if (kmajor) {
    cuuint64_t dims[3] = {(cuuint64_t)K, (cuuint64_t)rows_or_cols, (cuuint64_t)batch};
    cuuint64_t strides[2] = {(cuuint64_t)ld * 4, bstride ? bstride : 16};
    cuuint32_t box[3] = {32, (cuuint32_t)box_rows, 1};
    cuuint32_t es[3] = {1, 1, 1};
    r = enc(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, (void*)ptr, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
            CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  } else {
    cuuint64_t dims[3] = {(cuuint64_t)rows_or_cols, (cuuint64_t)K, (cuuint64_t)batch};
    cuuint64_t strides[2] = {(cuuint64_t)ld * 4, bstride ? bstride : 16};
    cuuint32_t box[3] = {32, 32, 1};
    cuuint32_t es[3] = {1, 1, 1};
    r = enc(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, (void*)ptr, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
            CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  }
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled failed (%d): kmajor=%d rows=%lld K=%lld ld=%lld sz=%lld batch=%lld", (int)r, (int)kmajor,
              (long long)rows_or_cols, (long long)K, (long long)ld, (long long)sz, (long long)batch);
    return LIP_ERR_CUDA;
  }
  return LIP_OK;
}

template <int BN, bool A_K, bool B_K, int CL>
int launch_tc(const TcGemmProblem& g, cudaStream_t st) {
  CUtensorMap maps[8];
  const TcOperand* ops[4] = {&g.A1, &g.B1, &g.A2, &g.B2};
  const int batched[4] = {g.a_batched, g.b_batched, g.a2_batched, g.b2_batched};
  const bool dual = g.A2.hi != nullptr;
  for (int i = 0; i < 4; ++i) {
    const TcOperand& o = *ops[(i >= 2 && !dual) ? i - 2 : i];
    const int bt = batched[(i >= 2 && !dual) ? i - 2 : i];
    const bool is_a = (i % 2 == 0);
    const bool km = is_a ? A_K : B_K;
    const int64_t K = (i >= 2 && dual) ? g.K2 : g.K;
    const int64_t rows = is_a ? g.M : g.N;
    const int box_rows = (is_a ? TBM : BN) / (CL == 1 ? 1 : 2);   // cluster mode: each CTA fetches half a tile
    int rc = make_map(&maps[2 * i], o.hi, km, rows, K, o.ld, o.sz, bt ? g.batch : 1, box_rows);
    if (rc) return rc;
    rc = make_map(&maps[2 * i + 1], o.lo, km, rows, K, o.ld, o.sz, bt ? g.batch : 1, box_rows);
    if (rc) return rc;
  }
  TcParams p;
  p.M = (int)g.M; p.N = (int)g.N; p.K1 = (int)g.K; p.K2 = dual ? (int)g.K2 : 0; p.batch = (int)g.batch;
  p.dbg = g_tc_dbg;
  p.kc = getenv("LIP_TC_KC") ? atoi(getenv("LIP_TC_KC")) : KC;
  p.merge = getenv("LIP_TC_MERGE") ? atoi(getenv("LIP_TC_MERGE")) : 0;
  p.a1_batched = g.a_batched; p.b1_batched = g.b_batched; p.a2_batched = g.a2_batched; p.b2_batched = g.b2_batched;
  p.C = g.C; p.C_lo = g.C_lo; p.c_sz = g.c_sz; p.c_sm = g.c_sm;
  p.scale = g.epi.scale;
  p.bias = g.epi.bias; p.bias_sz = g.epi.bias_sz;
  p.mask = g.epi.mask; p.mask_sm = g.epi.mask_sm;
  p.add = g.epi.add; p.add_sz = g.epi.add_sz; p.add_scale = g.epi.add_scale;
  p.colsum = g.colsum; p.colsum_sz = g.colsum_sz; p.colsum_ld = g.colsum_ld;
  p.b1_lo_nz = g.B1.lo_nz;
  using SL = SmemLayout<BN>;
  auto kern = gemm_tc_kernel<BN, A_K, B_K, CL>;
  static bool attr_set = false;
  if (!attr_set) {
    LIP_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, SL::BYTES));
    attr_set = true;
  }
  static int num_sms = 0;
  if (num_sms == 0) {
    int dev = 0;
    LIP_CHECK_CUDA(cudaGetDevice(&dev));
    LIP_CHECK_CUDA(cudaDeviceGetAttribute(&num_sms, cudaDevAttrMultiProcessorCount, dev));
  }
  const int64_t mt = ceil_div(g.M, TBM), nt = ceil_div(g.N, BN);
  cudaLaunchConfig_t cfg = {};
  cfg.blockDim = dim3(TC_THREADS);
  cfg.dynamicSmemBytes = SL::BYTES;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  if (CL == 1) {
    const int64_t ntiles = mt * nt * g.batch;
    cfg.gridDim = dim3((unsigned)(ntiles < num_sms ? ntiles : num_sms));
    cfg.numAttrs = 0;
  } else {
    const int64_t nsuper = ceil_div(mt, 2) * ceil_div(nt, 2) * g.batch;
    const int64_t max_clusters = num_sms / CL;
    cfg.gridDim = dim3((unsigned)((nsuper < max_clusters ? nsuper : max_clusters) * CL));
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = CL; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
  }
  LIP_CHECK_CUDA(cudaLaunchKernelEx(&cfg, kern, maps[0], maps[1], maps[2], maps[3], maps[4], maps[5], maps[6], maps[7], p));
  LIP_LAUNCH_CHECK();
  return LIP_OK;
}

template <int BN2, bool A_K, bool B_K>
int launch_tc2(const TcGemmProblem& g, cudaStream_t st) {
  CUtensorMap maps[8];
  const TcOperand* ops[4] = {&g.A1, &g.B1, &g.A2, &g.B2};
  const int batched[4] = {g.a_batched, g.b_batched, g.a2_batched, g.b2_batched};
  const bool dual = g.A2.hi != nullptr;
  for (int i = 0; i < 4; ++i) {
    const TcOperand& o = *ops[(i >= 2 && !dual) ? i - 2 : i];
    const int bt = batched[(i >= 2 && !dual) ? i - 2 : i];
    const bool is_a = (i % 2 == 0);
    const bool km = is_a ? A_K : B_K;
    const int64_t K = (i >= 2 && dual) ? g.K2 : g.K;
    const int64_t rows = is_a ? g.M : g.N;
    const int box_rows = is_a ? TBM : BN2 / 2;        // each CTA of the pair stages half of the B tile
    int rc = make_map(&maps[2 * i], o.hi, km, rows, K, o.ld, o.sz, bt ? g.batch : 1, box_rows);
    if (rc) return rc;
    rc = make_map(&maps[2 * i + 1], o.lo, km, rows, K, o.ld, o.sz, bt ? g.batch : 1, box_rows);
    if (rc) return rc;
  }
  TcParams p;
  p.M = (int)g.M; p.N = (int)g.N; p.K1 = (int)g.K; p.K2 = dual ? (int)g.K2 : 0; p.batch = (int)g.batch;
  p.dbg = g_tc_dbg;
  p.kc = getenv("LIP_TC_KC") ? atoi(getenv("LIP_TC_KC")) : KC;
  p.merge = getenv("LIP_TC_MERGE") ? atoi(getenv("LIP_TC_MERGE")) : 0;
  p.a1_batched = g.a_batched; p.b1_batched = g.b_batched; p.a2_batched = g.a2_batched; p.b2_batched = g.b2_batched;
  p.C = g.C; p.C_lo = g.C_lo; p.c_sz = g.c_sz; p.c_sm = g.c_sm;
  p.scale = g.epi.scale;
  p.bias = g.epi.bias; p.bias_sz = g.epi.bias_sz;
  p.mask = g.epi.mask; p.mask_sm = g.epi.mask_sm;
  p.add = g.epi.add; p.add_sz = g.epi.add_sz; p.add_scale = g.epi.add_scale;
  p.colsum = g.colsum; p.colsum_sz = g.colsum_sz; p.colsum_ld = g.colsum_ld;
  p.b1_lo_nz = g.B1.lo_nz;
  using SL = SmemLayout2<BN2>;
  auto kern = gemm_tc2_kernel<BN2, A_K, B_K>;
  static bool attr_set = false;
  if (!attr_set) {
    LIP_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, SL::BYTES));
    attr_set = true;
  }
  static int num_sms = 0;
  if (num_sms == 0) {
    int dev = 0;
    LIP_CHECK_CUDA(cudaGetDevice(&dev));
    LIP_CHECK_CUDA(cudaDeviceGetAttribute(&num_sms, cudaDevAttrMultiProcessorCount, dev));
  }
  const int64_t npairs = ceil_div(g.M, 2 * TBM) * ceil_div(g.N, BN2) * g.batch;
  const int64_t max_clusters = num_sms / 2;
  cudaLaunchConfig_t cfg = {};
  cfg.blockDim = dim3(TC_THREADS);
  cfg.dynamicSmemBytes = SL::BYTES;
  cfg.stream = st;
  cfg.gridDim = dim3((unsigned)((npairs < max_clusters ? npairs : max_clusters) * 2));
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = 2; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  LIP_CHECK_CUDA(cudaLaunchKernelEx(&cfg, kern, maps[0], maps[1], maps[2], maps[3], maps[4], maps[5], maps[6], maps[7], p));
  LIP_LAUNCH_CHECK();
  return LIP_OK;
}

template <bool A_K, bool B_K>
int launch_tc2w(const TcGemmProblem& g, cudaStream_t st) {
  CUtensorMap maps[8];
  const TcOperand* ops[4] = {&g.A1, &g.B1, &g.A2, &g.B2};
  const int batched[4] = {g.a_batched, g.b_batched, g.a2_batched, g.b2_batched};
  const bool dual = g.A2.hi != nullptr;
  for (int i = 0; i < 4; ++i) {
    const TcOperand& o = *ops[(i >= 2 && !dual) ? i - 2 : i];
    const int bt = batched[(i >= 2 && !dual) ? i - 2 : i];
    const bool is_a = (i % 2 == 0);
    const bool km = is_a ? A_K : B_K;
    const int64_t K = (i >= 2 && dual) ? g.K2 : g.K;
    const int64_t rows = is_a ? g.M : g.N;
    const int box_rows = is_a ? TBM : WN / 2;
    int rc = make_map(&maps[2 * i], o.hi, km, rows, K, o.ld, o.sz, bt ? g.batch : 1, box_rows);
    if (rc) return rc;
    rc = make_map(&maps[2 * i + 1], o.lo, km, rows, K, o.ld, o.sz, bt ? g.batch : 1, box_rows);
    if (rc) return rc;
  }
  TcParams p;
  p.M = (int)g.M; p.N = (int)g.N; p.K1 = (int)g.K; p.K2 = dual ? (int)g.K2 : 0; p.batch = (int)g.batch;
  p.dbg = g_tc_dbg;
  p.kc = getenv("LIP_TC_KC") ? atoi(getenv("LIP_TC_KC")) : KC;
  p.merge = 0;
  p.a1_batched = g.a_batched; p.b1_batched = g.b_batched; p.a2_batched = g.a2_batched; p.b2_batched = g.b2_batched;
  p.C = g.C; p.C_lo = g.C_lo; p.c_sz = g.c_sz; p.c_sm = g.c_sm;
  p.scale = g.epi.scale;
  p.bias = g.epi.bias; p.bias_sz = g.epi.bias_sz;
  p.mask = g.epi.mask; p.mask_sm = g.epi.mask_sm;
  p.add = g.epi.add; p.add_sz = g.epi.add_sz; p.add_scale = g.epi.add_scale;
  p.colsum = g.colsum; p.colsum_sz = g.colsum_sz; p.colsum_ld = g.colsum_ld;
  p.b1_lo_nz = g.B1.lo_nz;
  using SL = SmemLayoutW;
  auto kern = gemm_tc2w_kernel<A_K, B_K>;
  static bool attr_set = false;
  if (!attr_set) {
    LIP_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, SL::BYTES));
    attr_set = true;
  }
  static int num_sms = 0;
  if (num_sms == 0) {
    int dev = 0;
    LIP_CHECK_CUDA(cudaGetDevice(&dev));
    LIP_CHECK_CUDA(cudaDeviceGetAttribute(&num_sms, cudaDevAttrMultiProcessorCount, dev));
  }
  const int64_t npairs = ceil_div(g.M, 2 * TBM) * ceil_div(g.N, WN) * g.batch;
  const int64_t max_clusters = num_sms / 2;
  cudaLaunchConfig_t cfg = {};
  cfg.blockDim = dim3(TC_THREADS);
  cfg.dynamicSmemBytes = SL::BYTES;
  cfg.stream = st;
  cfg.gridDim = dim3((unsigned)((npairs < max_clusters ? npairs : max_clusters) * 2));
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = 2; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  LIP_CHECK_CUDA(cudaLaunchKernelEx(&cfg, kern, maps[0], maps[1], maps[2], maps[3], maps[4], maps[5], maps[6], maps[7], p));
  LIP_LAUNCH_CHECK();
  return LIP_OK;
}

}  // namespace

EncodeTiledFn tc_get_encode() { return get_encode(); }
int tc_make_map(CUtensorMap* map, const float* ptr, bool kmajor, int64_t rows_or_cols, int64_t K, int64_t ld, int64_t sz,
                int64_t batch, int box_rows) {
  return make_map(map, ptr, kmajor, rows_or_cols, K, ld, sz, batch, box_rows);
}

bool tc_available() {
  static int cached = -1;
  if (cached < 0) {
    int dev = 0, major = 0, minor = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return false;
    cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev);
    cudaDeviceGetAttribute(&minor, cudaDevAttrComputeCapabilityMinor, dev);
    cached = (major == 10 && minor == 0 && get_encode() != nullptr) ? 1 : 0;
  }
  return cached == 1;
}

int gemm_tc(const TcGemmProblem& g, cudaStream_t st) {
  if (g.M <= 0 || g.N <= 0 || g.batch <= 0) return LIP_OK;
  LIP_REQUIRE(g.A1.hi && g.A1.lo && g.B1.hi && g.B1.lo && g.C, "gemm_tc: null operand");
  LIP_REQUIRE(g.epi.act < 0 && g.epi.dphi_out == nullptr, "gemm_tc: activation epilogue is SIMT-only");
  const bool a_k = g.A1.major_k != 0, b_k = g.B1.major_k != 0;
  if (g.A2.hi) LIP_REQUIRE((g.A2.major_k != 0) == a_k && (g.B2.major_k != 0) == b_k, "gemm_tc: second pair must share majors");
  // (the kernel template also has a 2x2-cluster TMA-multicast mode, CL == 4: correct on B200 but the lock-step coupling cost more
  // than the L2 saving, so it is not instantiated)
  // cta_group::2 policy: -1 (default) = CTA pairs for the JVP-type GEMMs (K-major A, MN-major B) whose M fills whole
  // 256-row pair tiles - measured faster there (profiles/r01_launches_*), slower for the short-K weight-gradient /
  // delta-backprop GEMMs; 0 = never; 1 = always (when M > 128)
  static const int two_cta = getenv("LIP_TC_2CTA") ? atoi(getenv("LIP_TC_2CTA")) : -1;
  static const bool verbose = getenv("LIP_TC_VERBOSE") != nullptr;
  if (verbose) fprintf(stderr, "[lip] gemm_tc M=%lld N=%lld K=%lld K2=%lld batch=%lld a_k=%d b_k=%d two_cta=%d\n", (long long)g.M,
                       (long long)g.N, (long long)g.K, (long long)g.K2, (long long)g.batch, (int)a_k, (int)b_k, two_cta);
  // Ragged M (e.g. the 784-row weight gradient of the first MNIST layer): run the rows that fill whole 256-row pair tiles
  // on the wide kernel and the remaining (< 256) rows as a second launch, instead of padding 784 -> 1024 or giving the whole
  // problem to the 128-wide kernel.
  static const int split_m = getenv("LIP_TC_SPLIT_M") ? atoi(getenv("LIP_TC_SPLIT_M")) : 1;
  if (split_m && g_tc_force2 < 0 && g.M > 2 * TBM && g.M % (2 * TBM) != 0 && g.N > 128 && !g.colsum &&
      (g.M / (2 * TBM)) * ceil_div(g.N, WN) * g.batch >= 74) {
    const int64_t M0 = g.M / (2 * TBM) * (2 * TBM);
    auto shift = [&](TcOperand o, bool batched_unused) {
      (void)batched_unused;
      if (o.hi) {
        const int64_t off = o.major_k ? M0 * o.ld : M0;     // K-major: rows are M; MN-major: M is the contiguous index
        o.hi += off; o.lo += off;
      }
      return o;
    };
    TcGemmProblem head = g, tail = g;
    head.M = M0;
    tail.M = g.M - M0;
    tail.A1 = shift(g.A1, g.a_batched);
    tail.A2 = shift(g.A2, g.a2_batched);
    tail.C = g.C + M0 * g.c_sm;
    if (g.C_lo) tail.C_lo = g.C_lo + M0 * g.c_sm;
    if (g.epi.mask) tail.epi.mask = g.epi.mask + M0 * g.epi.mask_sm;
    if (g.epi.add) tail.epi.add = g.epi.add + M0 * g.c_sm;
    int rc = gemm_tc(head, st);
    if (rc) return rc;
    return gemm_tc(tail, st);
  }
  // wide CTA-pair tiles (256 x 256; measured ~22 % faster per tile than 128-wide tiles, profiles/r01_gemm_microbench_wide.txt):
  // used when the padded tile area does not grow by more than 10 % over 128 x 128 tiles.  LIP_TC_WIDE: -1 auto (default),
  // 0 never, 1 whenever M > 128 and N > 128.
  static const int wide = getenv("LIP_TC_WIDE") ? atoi(getenv("LIP_TC_WIDE")) : -1;
  {
    const double area_w = (double)(ceil_div(g.M, 2 * TBM) * 2 * TBM) * (double)(ceil_div(g.N, WN) * WN);
    const double area_n = (double)(ceil_div(g.M, TBM) * TBM) * (double)(ceil_div(g.N, 128) * 128);
    // wide pair tiles only when there are enough of them to fill the 74 CTA pairs (small probe batches, e.g. the 4-probe
    // mat-vecs inside SLQ, get more parallelism from 128-wide tiles)
    const int64_t npairs_w = ceil_div(g.M, 2 * TBM) * ceil_div(g.N, WN) * g.batch;
    const bool fits = area_w <= 1.10 * area_n && npairs_w >= 74;
    const bool use_w = g_tc_force2 == 2 || (g_tc_force2 < 0 && (wide == 1 || (wide < 0 && fits)) && g.M > TBM && g.N > 128);
    if (use_w) {
      if (a_k && !b_k) return launch_tc2w<true, false>(g, st);
      if (!a_k && !b_k) return launch_tc2w<false, false>(g, st);
      if (a_k && b_k) return launch_tc2w<true, true>(g, st);
    }
  }
  const bool auto2 = a_k && !b_k && (g.M % (2 * TBM) == 0) && ceil_div(g.M, 2 * TBM) * ceil_div(g.N, 128) * g.batch >= 74;
  const bool use2 = g_tc_force2 >= 0 ? (g_tc_force2 == 1) : (two_cta < 0 ? auto2 : two_cta != 0);
  if (use2 && g.M > TBM) {
    if (a_k && !b_k) return launch_tc2<128, true, false>(g, st);
    if (!a_k && !b_k) return launch_tc2<128, false, false>(g, st);
    if (a_k && b_k) return launch_tc2<128, true, true>(g, st);
  }
  if (a_k && !b_k) return launch_tc<128, true, false, 1>(g, st);
  if (!a_k && !b_k) return launch_tc<128, false, false, 1>(g, st);
  if (a_k && b_k) return launch_tc<128, true, true, 1>(g, st);
  set_error("gemm_tc: unsupported operand majors (A MN-major with B K-major)");
  return LIP_ERR_INVALID;
}

int tf32_split3(const float* src, int64_t sz_src, int64_t ld_src, float* hi, float* lo, int64_t sz_dst, int64_t ld_dst,
                int64_t batch, int64_t rows, int64_t cols, cudaStream_t st, int* lo_nz) {
  if (rows <= 0 || cols <= 0 || batch <= 0) return LIP_OK;
  if (ld_src == cols && ld_dst == cols && batch <= 65535) {
    const int64_t n = rows * cols;
    int64_t g = ceil_div(n, 256 * 8);
    const int64_t cap = ceil_div((int64_t)148 * 16, batch);      // ~16 resident CTAs' worth of blocks per SM overall
    if (g > cap) g = cap;
    if (g < 1) g = 1;
    dim3 grid((unsigned)g, (unsigned)batch);
    tf32_split_flat_kernel<<<grid, 256, 0, st>>>(src, sz_src, hi, lo, sz_dst, n, lo_nz);
    LIP_LAUNCH_CHECK();
    return LIP_OK;
  }
  int64_t gx = ceil_div(ld_dst, 256);
  if (gx > 64) gx = 64;
  for (int64_t z0 = 0; z0 < batch; z0 += 65535) {
    const int64_t zc = batch - z0 < 65535 ? batch - z0 : 65535;
    for (int64_t r0 = 0; r0 < rows; r0 += 65535) {
      const int64_t rc = rows - r0 < 65535 ? rows - r0 : 65535;
      dim3 grid((unsigned)gx, (unsigned)rc, (unsigned)zc);
      tf32_split_kernel<<<grid, 256, 0, st>>>(src + z0 * sz_src + r0 * ld_src, sz_src, ld_src, hi + z0 * sz_dst + r0 * ld_dst,
                                              lo + z0 * sz_dst + r0 * ld_dst, sz_dst, ld_dst, cols, lo_nz);
      LIP_LAUNCH_CHECK();
    }
  }
  return LIP_OK;
}

int tf32_split(const float* src, int64_t ld_src, float* hi, float* lo, int64_t ld_dst, int64_t rows, int64_t cols,
               cudaStream_t st) {
  return tf32_split3(src, 0, ld_src, hi, lo, 0, ld_dst, 1, rows, cols, st, nullptr);
}

}  // namespace lip

// ---- self test: tensor-core GEMM vs the exact SIMT GEMM on random data ---------------------------------------
namespace {
__global__ void fill_random_kernel(float* x, long long n, unsigned seed, float scale) {
  long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (i >= n) return;
  unsigned s = (unsigned)(i * 2654435761u) ^ seed;
  s ^= s >> 16; s *= 0x7feb352du; s ^= s >> 15; s *= 0x846ca68bu; s ^= s >> 16;
  x[i] = scale * ((float)(s & 0xFFFFFF) / 8388608.f - 1.f);
}
__global__ void max_rel_err_kernel(const float* a, const float* b, long long n, float* num, float* den) {
  __shared__ float sn[256], sd[256];
  float ln = 0.f, ld = 0.f;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    float d = a[i] - b[i];
    ln += d * d;
    ld += b[i] * b[i];
  }
  sn[threadIdx.x] = ln; sd[threadIdx.x] = ld;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if (threadIdx.x < o) { sn[threadIdx.x] += sn[threadIdx.x + o]; sd[threadIdx.x] += sd[threadIdx.x + o]; }
    __syncthreads();
  }
  if (threadIdx.x == 0) { atomicAdd(num, sn[0]); atomicAdd(den, sd[0]); }
}
}  // namespace

extern "C" int lip_bench_tc_gemm(int32_t variant, int64_t M, int64_t N, int64_t K, int64_t batch, int32_t iters, int32_t dbg,
                                 int32_t two_cta, float* ms_per_iter, lip_stream_t stream) {
  using namespace lip;
  LIP_REQUIRE(variant >= 0 && variant <= 2 && M > 0 && N > 0 && K > 0 && batch > 0 && iters > 0 && ms_per_iter, "bench: bad argument");
  if (!tc_available()) { set_error("bench: tcgen05 path unavailable"); return LIP_ERR_UNSUPPORTED; }
  cudaStream_t st = (cudaStream_t)stream;
  auto pad = [](int64_t x) { return (x + 31) / 32 * 32; };
  const bool a_k = (variant != 1), b_k = (variant == 2);
  const bool a_batched = (variant == 2), b_batched = (variant != 2);
  const int64_t a_rows = a_k ? M : K, a_cols = a_k ? K : M, b_rows = b_k ? N : K, b_cols = b_k ? K : N;
  const int64_t lda = pad(a_cols), ldb = pad(b_cols), a_sz = a_rows * lda, b_sz = b_rows * ldb;
  const int64_t na = (a_batched ? batch : 1) * a_sz + 64, nb = (b_batched ? batch : 1) * b_sz + 64, nc = batch * M * N;
  float *Ah, *Al, *Bh, *Bl, *C;
  LIP_CHECK_CUDA(cudaMalloc(&Ah, 4 * na)); LIP_CHECK_CUDA(cudaMalloc(&Al, 4 * na));
  LIP_CHECK_CUDA(cudaMalloc(&Bh, 4 * nb)); LIP_CHECK_CUDA(cudaMalloc(&Bl, 4 * nb)); LIP_CHECK_CUDA(cudaMalloc(&C, 4 * nc));
  if (getenv("LIP_BENCH_RANDOM")) {   // realistic switching activity (power / clocks) instead of all-zero operands
    fill_random_kernel<<<(unsigned)ceil_div(na, 256), 256, 0, st>>>(Ah, na, 11u, 1.f);
    fill_random_kernel<<<(unsigned)ceil_div(nb, 256), 256, 0, st>>>(Bh, nb, 23u, 1.f);
    tf32_split(Ah, lda, Ah, Al, lda, (a_batched ? batch : 1) * a_rows, lda, st);
    tf32_split(Bh, ldb, Bh, Bl, ldb, (b_batched ? batch : 1) * b_rows, ldb, st);
  } else {
    cudaMemsetAsync(Ah, 0, 4 * na, st); cudaMemsetAsync(Al, 0, 4 * na, st);
    cudaMemsetAsync(Bh, 0, 4 * nb, st); cudaMemsetAsync(Bl, 0, 4 * nb, st);
  }
  TcGemmProblem tp;
  tp.M = M; tp.N = N; tp.K = K; tp.batch = batch;
  tp.A1.hi = Ah; tp.A1.lo = Al; tp.A1.sz = a_sz; tp.A1.ld = lda; tp.A1.major_k = a_k;
  tp.B1.hi = Bh; tp.B1.lo = Bl; tp.B1.sz = b_sz; tp.B1.ld = ldb; tp.B1.major_k = b_k;
  tp.a_batched = a_batched; tp.b_batched = b_batched;
  tp.C = C; tp.c_sz = M * N; tp.c_sm = N;
  g_tc_dbg = dbg; g_tc_force2 = two_cta;
  int rc = gemm_tc(tp, st);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  if (!rc) {
    cudaEventRecord(e0, st);
    for (int i = 0; i < iters && !rc; ++i) rc = gemm_tc(tp, st);
    cudaEventRecord(e1, st);
    cudaError_t e = cudaEventSynchronize(e1);
    if (e != cudaSuccess) { set_error("bench: %s", cudaGetErrorString(e)); rc = LIP_ERR_CUDA; }
    float ms = 0.f;
    cudaEventElapsedTime(&ms, e0, e1);
    *ms_per_iter = ms / iters;
  }
  g_tc_dbg = 0; g_tc_force2 = -1;
  cudaEventDestroy(e0); cudaEventDestroy(e1);
  cudaFree(Ah); cudaFree(Al); cudaFree(Bh); cudaFree(Bl); cudaFree(C);
  return rc;
}

extern "C" int lip_selftest_tc_gemm(int32_t variant, int64_t M, int64_t N, int64_t K, int64_t batch, float* max_rel_err,
                                    lip_stream_t stream) {
  using namespace lip;
  LIP_REQUIRE(variant >= 0 && variant <= 2 && M > 0 && N > 0 && K > 0 && batch > 0 && max_rel_err, "selftest: bad argument");
  if (!tc_available()) { set_error("selftest: tcgen05 path unavailable on this device"); return LIP_ERR_UNSUPPORTED; }
  cudaStream_t st = (cudaStream_t)stream;
  auto pad = [](int64_t x) { return (x + 31) / 32 * 32; };
  // logical operands: A [M x K], B [K x N] per batch (A shared for variant 0/1, B shared for variant 2)
  const bool a_k = (variant != 1), b_k = (variant == 2);
  const bool a_batched = (variant == 2), b_batched = (variant != 2);
  const int64_t a_rows = a_k ? M : K, a_cols = a_k ? K : M;   // memory layout rows x cols (cols contiguous)
  const int64_t b_rows = b_k ? N : K, b_cols = b_k ? K : N;
  const int64_t lda = pad(a_cols), ldb = pad(b_cols);
  const int64_t a_sz = a_rows * lda, b_sz = b_rows * ldb;
  const int64_t na = (a_batched ? batch : 1) * a_sz + 64, nb = (b_batched ? batch : 1) * b_sz + 64;
  float *A, *Ah, *Al, *Bm, *Bh, *Bl, *C0, *C1, *mask, *stats;
  LIP_CHECK_CUDA(cudaMalloc(&A, 4 * na)); LIP_CHECK_CUDA(cudaMalloc(&Ah, 4 * na)); LIP_CHECK_CUDA(cudaMalloc(&Al, 4 * na));
  LIP_CHECK_CUDA(cudaMalloc(&Bm, 4 * nb)); LIP_CHECK_CUDA(cudaMalloc(&Bh, 4 * nb)); LIP_CHECK_CUDA(cudaMalloc(&Bl, 4 * nb));
  const int64_t nc = batch * M * N;
  LIP_CHECK_CUDA(cudaMalloc(&C0, 4 * nc)); LIP_CHECK_CUDA(cudaMalloc(&C1, 4 * nc));
  LIP_CHECK_CUDA(cudaMalloc(&mask, 4 * M * N)); LIP_CHECK_CUDA(cudaMalloc(&stats, 8));
  fill_random_kernel<<<(unsigned)ceil_div(na, 256), 256, 0, st>>>(A, na, 11u, 1.f);
  fill_random_kernel<<<(unsigned)ceil_div(nb, 256), 256, 0, st>>>(Bm, nb, 23u, 1.f);
  fill_random_kernel<<<(unsigned)ceil_div(M * N, 256), 256, 0, st>>>(mask, M * N, 37u, 1.f);
  int rc = tf32_split(A, lda, Ah, Al, lda, (a_batched ? batch : 1) * a_rows, lda, st);
  if (!rc) rc = tf32_split(Bm, ldb, Bh, Bl, ldb, (b_batched ? batch : 1) * b_rows, ldb, st);
  GemmProblem sp;
  sp.M = M; sp.N = N; sp.K = K; sp.batch = batch;
  sp.A1 = {A, a_batched ? a_sz : 0, a_k ? lda : 1, a_k ? 1 : lda};
  sp.B1 = {Bm, b_batched ? b_sz : 0, b_k ? 1 : ldb, b_k ? ldb : 1};
  sp.C = C0; sp.c_sz = M * N; sp.c_sm = N;
  sp.epi.scale = 0.5f; sp.epi.mask = mask; sp.epi.mask_sm = N;
  if (!rc) rc = gemm_simt(sp, st);
  TcGemmProblem tp;
  tp.M = M; tp.N = N; tp.K = K; tp.batch = batch;
  tp.A1.hi = Ah; tp.A1.lo = Al; tp.A1.sz = a_sz; tp.A1.ld = lda; tp.A1.major_k = a_k;
  tp.B1.hi = Bh; tp.B1.lo = Bl; tp.B1.sz = b_sz; tp.B1.ld = ldb; tp.B1.major_k = b_k;
  tp.a_batched = a_batched; tp.b_batched = b_batched;
  tp.C = C1; tp.c_sz = M * N; tp.c_sm = N;
  tp.epi = sp.epi;
  if (!rc) rc = gemm_tc(tp, st);
  float h[2] = {0.f, 0.f};
  if (!rc && getenv("LIP_TC_DEBUG")) {
    // structured probe: A = 1 for k == kd (else 0), B[k][n] = n + 1000 k  ->  C[m][n] = 0.5 * mask * (n + 1000 kd)
    cudaStreamSynchronize(st);
    std::vector<float> c0(64), c1(64);
    for (int row : {0, 1, 33}) {
      cudaMemcpy(c0.data(), C0 + (size_t)row * N, 4 * 8, cudaMemcpyDeviceToHost);
      cudaMemcpy(c1.data(), C1 + (size_t)row * N, 4 * 8, cudaMemcpyDeviceToHost);
      fprintf(stderr, "  row %d simt:", row);
      for (int i = 0; i < 8; ++i) fprintf(stderr, " %9.4f", c0[i]);
      fprintf(stderr, "\n  row %d tc  :", row);
      for (int i = 0; i < 8; ++i) fprintf(stderr, " %9.4f", c1[i]);
      fprintf(stderr, "\n");
    }
  }
  if (!rc) {
    cudaMemsetAsync(stats, 0, 8, st);
    max_rel_err_kernel<<<256, 256, 0, st>>>(C1, C0, nc, stats, stats + 1);
    cudaError_t e = cudaMemcpyAsync(h, stats, 8, cudaMemcpyDeviceToHost, st);
    if (e == cudaSuccess) e = cudaStreamSynchronize(st);
    if (e != cudaSuccess) { set_error("selftest: %s", cudaGetErrorString(e)); rc = LIP_ERR_CUDA; }
  }
  cudaFree(A); cudaFree(Ah); cudaFree(Al); cudaFree(Bm); cudaFree(Bh); cudaFree(Bl); cudaFree(C0); cudaFree(C1);
  cudaFree(mask); cudaFree(stats);
  if (rc) return rc;
  *max_rel_err = h[1] > 0.f ? sqrtf(h[0] / h[1]) : (h[0] > 0.f ? 1e30f : 0.f);
  return LIP_OK;
}
