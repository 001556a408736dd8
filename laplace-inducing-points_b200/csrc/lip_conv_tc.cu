// lip_conv_tc.cu — tcgen05 implicit-GEMM convolutions (3xTF32 fp32 emulation) for the residual conv programs (lip_resnet.cu).
//
// The three conv GEMMs of a conv + BatchNorm unit (JVP, kernel gradient, delta back-propagation; lip_resnet.cu header) on
// the tensor cores WITHOUT a patch buffer: the A operand is staged by TMA straight from the NHWC image through a 4-D tensor
// map {C, W, H, images}.  With C a multiple of 32, one 32-float k-block of the patch matrix (column order (dy, dx, c), the row
// order of a flax HWIO kernel) is ONE filter tap and 32 channels = one 128-byte swizzle row per pixel, and a tile of 128
// consecutive output pixels is a box {32 c, bw, bh, bn} of whole image rows (bw*bh*bn = 128).  The tap only shifts the box
// origin by (dx - pad, dy - pad); TMA's out-of-bounds zero fill IS the zero padding of the convolution.
//   conv_tc_kernel<NB, 1>   A K-major : rows = pixels, K = (tap, c)     JVP  dH[b] = patches(X).dW[b] + patches(T[b]).W   (dual K)
//                                                                        back-prop  cot[b] = patches'(Dh[b]).Wt   (shift = pad - d)
//   conv_tc_kernel<NB, 2>   A MN-major: rows = (tap, c), K = pixels      kernel gradient  gW[b] = patches(X)^T.Dh[b], split-K with a
//                                                                        deterministic second-pass reduction (conv_splitk_reduce)
// B is always MN-major ([K, cout] row-major: tangent / bound kernels, deltas) and NB = cout tile width (32, 64 or 128: the UMMA N).
// Pipeline, warp roles, chunked TMEM drain (fp32 TMEM accumulation truncates) and the fused epilogue are those of gemm_tc_kernel
// (lip_gemm_tc.cu); the shared device code lives in lip_tc_dev.cuh.  Stride 2 is the same box traversed with TMA element
// strides {1, 2, 2, 1} (forward / kernel gradient); the delta back-propagation of a strided conv runs as the stride-1 transposed
// conv of the zero-upsampled delta image (conv_tc_upsample2).  1x1 / 3x3 convs with channel counts that are multiples of 32 and
// whose output rows tile the 128-pixel box are eligible (conv_tc_supported); the rest (the 3-channel stem) stays on the SIMT
// implicit GEMM.
#include <cuda.h>

#include <stdlib.h>
#include <vector>

#include "lip_tc_dev.cuh"
#include "lip_conv_tc.cuh"

namespace lip {

namespace {

struct ConvGeo {
  int W, H, P;               // OUTPUT pixel grid of the conv (rows of the patch matrix); the gathered image is stride x larger
  int stride;                // 1 or 2 (TMA traversal stride of the box over the gathered image)
  int C, cblocks;            // channels of the gathered image, C / 32
  int kw;                    // taps per kernel row
  int sgn, off_h, off_w;     // tap (dy, dx) shifts the box by (sgn*dy + off_h, sgn*dx + off_w)
  int imgs;                  // images per batch entry (points M)
  int Kc;                    // AMODE 2: patch columns = taps * C (rows of the output)
  int ksplit, kb_per;        // AMODE 2: K slices and k-blocks per slice
  int merge;                 // 1: one N = 2*NB MMA computes A_hi x [B_hi | B_lo] (see the MMA issuer); 0: three N = NB MMAs
  int fold, cpb;             // fold > 1: a 128-column tile = `fold` probes x N columns (cpb = N / 32 chunks each); A is shared
};

__device__ __forceinline__ void tma_load_4d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1, int c2, int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
      ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}

// Fused BatchNorm-JVP epilogue (ConvBnEpilogue in lip_conv_tc.cuh); same staging / coalescing scheme as tile_epilogue.
template <int NBLK, int RB>
__device__ __forceinline__ void bn_tile_epilogue(const TcParams& p, const ConvBnEpilogue& e, float (&acc)[32 * NBLK], float* stg,
                                                 int m0, int n0, int z, int q, int h, int lane) {
  const int mrow0 = m0 + q * 32;
  int nrows = p.M - mrow0;
  nrows = nrows > 32 ? 32 : nrows;
  const long long i0 = (long long)mrow0 * p.c_sm;              // offset inside one batch entry
  const long long zc = (long long)z * p.c_sz + i0;
#pragma unroll
  for (int cc = 0; cc < NBLK; ++cc) {
    const int n = n0 + h * (32 * NBLK) + cc * 32 + lane;
#pragma unroll
    for (int i = 0; i < 32; ++i) stg[lane * STG_LD + i] = acc[cc * 32 + i];
    __syncwarp();
    if (n < p.N && nrows > 0) {
      const float gv = __ldg(e.g + n);
      const float ds = __ldg(e.dscale + (long long)z * e.pstride + n), db = __ldg(e.dbeta + (long long)z * e.pstride + n);
      const float* xp = e.xhat + i0 + n;
      const float* mp = e.mask ? e.mask + i0 + n : nullptr;
      const float* sh = e.skip_hi ? e.skip_hi + zc + n : nullptr;
      const float* sl = e.skip_hi ? e.skip_lo + zc + n : nullptr;
      const float* pp = e.pre ? e.pre + zc + n : nullptr;
      float* ch = p.C + zc + n;
      float* cl = p.C_lo + zc + n;
      const float* sp = stg + lane;
      const long long cs = p.c_sm;
      int r0 = 0;
      for (; r0 + RB <= nrows; r0 += RB) {
        float xv[RB], mv[RB], s1[RB], s2[RB], pv[RB];
#pragma unroll
        for (int r = 0; r < RB; ++r) {
          pv[r] = pp ? __ldg(pp + (r0 + r) * cs) : 0.f;
          xv[r] = __ldg(xp + (r0 + r) * cs);
          mv[r] = mp ? __ldg(mp + (r0 + r) * cs) : 1.f;
          s1[r] = sh ? __ldg(sh + (r0 + r) * cs) : 0.f;
          s2[r] = sh ? __ldg(sl + (r0 + r) * cs) : 0.f;
        }
#pragma unroll
        for (int r = 0; r < RB; ++r) {
          float v = fmaf(gv, sp[(r0 + r) * STG_LD] + pv[r], fmaf(xv[r], ds, db));
          v = (v + (s1[r] + s2[r])) * mv[r];
          const float hh = tf32_rna(v);
          ch[(r0 + r) * cs] = hh;
          cl[(r0 + r) * cs] = tf32_rna(v - hh);
        }
      }
      for (; r0 < nrows; ++r0) {
        float v = fmaf(gv, sp[r0 * STG_LD] + (pp ? __ldg(pp + r0 * cs) : 0.f), fmaf(__ldg(xp + r0 * cs), ds, db));
        if (sh) v += __ldg(sh + r0 * cs) + __ldg(sl + r0 * cs);
        if (mp) v *= __ldg(mp + r0 * cs);
        const float hh = tf32_rna(v);
        ch[r0 * cs] = hh;
        cl[r0 * cs] = tf32_rna(v - hh);
      }
    }
    __syncwarp();
  }
}

template <int NB>
struct ConvSmem {
  static constexpr int A_TILE = TBM * TBK * 4;   // 16 KB
  static constexpr int B_TILE = NB * TBK * 4;    // NB/32 chunks of 4 KB
  static constexpr int STAGE = 2 * A_TILE + 2 * B_TILE;
  static constexpr int STAGES = (NB == 128) ? 3 : 4;
  static constexpr int STAGING = 8 * 32 * STG_LD * 4;
  static constexpr int BYTES = STAGES * STAGE + STAGING + 1024 + 256;
};

// EPI: 0 = plain fused epilogue (tile_epilogue), 1 = BatchNorm-JVP epilogue (AMODE 1), 2 = probes folded into the tile width
template <int NB, int AMODE, int EPI>
__global__ void __launch_bounds__(TC_THREADS, 1)
conv_tc_kernel(const __grid_constant__ CUtensorMap mA1h, const __grid_constant__ CUtensorMap mA1l,
               const __grid_constant__ CUtensorMap mB1h, const __grid_constant__ CUtensorMap mB1l,
               const __grid_constant__ CUtensorMap mA2h, const __grid_constant__ CUtensorMap mA2l,
               const __grid_constant__ CUtensorMap mB2h, const __grid_constant__ CUtensorMap mB2l, TcParams p0, ConvGeo g,
               ConvBnEpilogue bn) {
  using SL = ConvSmem<NB>;
  constexpr bool A_K = (AMODE == 1);
  constexpr int TMEM_COLS = 512;          // 2 buffers x (cross-term tile + main tile) x 128 columns (NB of them used)
  constexpr int TSTRIDE = 128;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t stg_base = smem_base + SL::STAGES * SL::STAGE;
  const uint32_t bar_base = stg_base + SL::STAGING;
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (SL::STAGES + s); };
  auto tfull_bar = [&](int b) { return bar_base + 8u * (2 * SL::STAGES + b); };
  auto tempty_bar = [&](int b) { return bar_base + 8u * (2 * SL::STAGES + 2 + b); };
  const uint32_t tmem_slot = bar_base + 8u * (2 * SL::STAGES + 4);

  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);
  const int lane = threadIdx.x & 31;
  const int nk1 = (p0.K1 + TBK - 1) / TBK, nk2 = (p0.K2 + TBK - 1) / TBK;
  const int nk = nk1 + nk2;
  const bool skip_b1lo = p0.b1_lo_nz != nullptr && *reinterpret_cast<const volatile int*>(p0.b1_lo_nz) == 0;
  const int KCr = p0.kc;
  const int S = (AMODE == 2) ? g.ksplit : 1;
  const int G = (EPI == 2) ? g.fold : 1;      // probes folded into the N dimension of one tile (1: none)
  const int mt = (p0.M + TBM - 1) / TBM, nt = G > 1 ? 1 : (p0.N + NB - 1) / NB;
  const int nz = G > 1 ? (p0.batch + G - 1) / G : p0.batch;      // batch entries, or groups of G probes
  const long long ntiles = (long long)mt * nt * nz * S;
  // tile t -> (m tile, n tile, batch entry or probe group z, K slice s): m fastest, slices of one output tile far apart
  auto tile_coords = [&](long long t, int& m0, int& n0, int& z, int& s, int& kb_lo, int& kb_hi) {
    const int tm = (int)(t % mt);
    long long r = t / mt;
    const int tn = (int)(r % nt);
    r /= nt;
    z = (int)(r % nz);
    s = (int)(r / nz);
    m0 = tm * TBM; n0 = tn * NB;
    if (AMODE == 2) {
      kb_lo = s * g.kb_per;
      kb_hi = kb_lo + g.kb_per < nk ? kb_lo + g.kb_per : nk;
    } else {
      kb_lo = 0; kb_hi = nk;
    }
  };

  if (warp == 0 && lane == 0) {
    for (int s = 0; s < SL::STAGES; ++s) { mbar_init(full_bar(s), 1); mbar_init(empty_bar(s), 1); }
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
  tc_fence_after();
  uint32_t tmem_base;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot));

  if (warp == 0) {
    // ================= TMA producer =================
    uint32_t it = 0;
    for (long long t = blockIdx.x; t < ntiles; t += gridDim.x) {
      int m0, n0, z, ks, kb_lo, kb_hi;
      tile_coords(t, m0, n0, z, ks, kb_lo, kb_hi);
      // AMODE 1: the tile's 128 pixel rows start at (img0, h0, w0) of batch entry z
      int img0 = 0, h0 = 0, w0 = 0;
      // AMODE 2: the four 32-column chunks of the A tile are four (tap, channel block) pairs
      int ch_c[4], ch_w[4], ch_h[4];
      if (AMODE == 1) {
        img0 = m0 / g.P;
        const int rem = m0 - img0 * g.P;
        h0 = rem / g.W;
        w0 = rem - h0 * g.W;
      } else {
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          const int mc = m0 + 32 * c;
          const int tap = mc / g.C;
          const int dy = tap / g.kw, dx = tap - dy * g.kw;
          ch_c[c] = mc < g.Kc ? mc - tap * g.C : g.C;       // beyond the last patch column: a fully out-of-bounds (zero) box
          ch_w[c] = g.sgn * dx + g.off_w;
          ch_h[c] = g.sgn * dy + g.off_h;
        }
      }
      for (int kb = kb_lo; kb < kb_hi; ++kb, ++it) {
        const int s = it % SL::STAGES;
        const uint32_t ph = (it / SL::STAGES) & 1;
        mbar_wait(empty_bar(s), ph ^ 1);
        const uint32_t st = smem_base + s * SL::STAGE;
        if (elect_one()) {
          const bool second = kb >= nk1;
          const bool no_blo = skip_b1lo && !second;
          mbar_arrive_expect_tx(full_bar(s), SL::STAGE - (no_blo ? SL::B_TILE : 0));
          const int kk = second ? kb - nk1 : kb;
          const CUtensorMap* ah = second ? &mA2h : &mA1h;
          const CUtensorMap* al = second ? &mA2l : &mA1l;
          const CUtensorMap* bh = second ? &mB2h : &mB1h;
          const CUtensorMap* bl = second ? &mB2l : &mB1l;
          const int za = (second ? p0.a2_batched : p0.a1_batched) ? z : 0;      // (fold mode: A is shared)
          const int zb = (second ? p0.b2_batched : p0.b1_batched) ? z : 0;
          if (AMODE == 1) {
            const int tap = kk / g.cblocks, cb = kk - tap * g.cblocks;
            const int dy = tap / g.kw, dx = tap - dy * g.kw;
            const int cw = w0 * g.stride + g.sgn * dx + g.off_w, chh = h0 * g.stride + g.sgn * dy + g.off_h;
            const int cn = za * g.imgs + img0;
            tma_load_4d(st, ah, full_bar(s), cb * 32, cw, chh, cn);
            tma_load_4d(st + SL::A_TILE, al, full_bar(s), cb * 32, cw, chh, cn);
          } else {
            const int k0 = kk * TBK;                    // first of this k-block's 32 pixel rows
            const int img = k0 / g.P;
            const int rem = k0 - img * g.P;
            const int ph0 = rem / g.W, pw0 = rem - ph0 * g.W;
            const int cn = za * g.imgs + img;
#pragma unroll
            for (int c = 0; c < 4; ++c) {
              tma_load_4d(st + c * (TBK * 128), ah, full_bar(s), ch_c[c], pw0 * g.stride + ch_w[c], ph0 * g.stride + ch_h[c], cn);
              tma_load_4d(st + SL::A_TILE + c * (TBK * 128), al, full_bar(s), ch_c[c], pw0 * g.stride + ch_w[c],
                          ph0 * g.stride + ch_h[c], cn);
            }
          }
          const int k0b = kk * TBK;
#pragma unroll
          for (int c = 0; c < NB / 32; ++c) {
            // fold mode: chunk c = 32 columns of probe z*G + c/cpb (a probe index past the batch is an all-zero box)
            const int bn0 = G > 1 ? (c % g.cpb) * 32 : n0 + 32 * c;
            const int bz = G > 1 ? z * G + c / g.cpb : zb;
            tma_load_3d(st + 2 * SL::A_TILE + c * (TBK * 128), bh, full_bar(s), bn0, k0b, bz);
            if (!no_blo) tma_load_3d(st + 2 * SL::A_TILE + SL::B_TILE + c * (TBK * 128), bl, full_bar(s), bn0, k0b, bz);
          }
        }
        __syncwarp();
      }
    }
  } else if (warp == 1) {
    // ================= MMA issuer =================
    // The B tile is MN-major: 32-column chunks 4 KB apart, and its lo chunks directly follow its hi chunks.  One N = 2*NB MMA
    // therefore computes A_hi x [B_hi | B_lo]: main term in TMEM columns [0, NB), first cross term in [NB, 2NB); the second
    // cross term A_lo x B_hi accumulates on top of the first.  A_hi is read from shared memory once instead of twice - the
    // narrow-N kernels are bound by exactly those reads.
    constexpr uint32_t idesc = make_idesc(TBM, NB, !A_K, true);
    constexpr uint32_t idesc2 = make_idesc(TBM, 2 * NB, !A_K, true);
    constexpr uint32_t A_LBO = A_K ? 16 : TBK * 128, B_LBO = TBK * 128;
    constexpr uint32_t A_SBO = A_K ? 1024 : 512, B_SBO = 512;
    constexpr uint32_t A_LT = A_K ? 2 : 1, B_LT = 1;
    constexpr uint32_t A_KSTEP = A_K ? 32 : 1024, B_KSTEP = 1024;
    uint32_t it = 0, ck = 0;
    for (long long t = blockIdx.x; t < ntiles; t += gridDim.x) {
      int m0, n0, z, ks, kb_lo, kb_hi;
      tile_coords(t, m0, n0, z, ks, kb_lo, kb_hi);
      const int nkt = kb_hi - kb_lo;
      const int nchunks = (nkt + KCr - 1) / KCr;
      for (int c = 0; c < nchunks; ++c, ++ck) {
        const uint32_t buf = ck & 1, cph = (ck >> 1) & 1;
        mbar_wait(tempty_bar(buf), cph ^ 1);
        tc_fence_after();
        const uint32_t t_main = tmem_base + buf * (2 * TSTRIDE), t_cross = t_main + (g.merge == 2 ? TSTRIDE : NB);
        const int kq_end = (c + 1) * KCr < nkt ? (c + 1) * KCr : nkt;
        for (int kq = c * KCr; kq < kq_end; ++kq, ++it) {
          const int s = it % SL::STAGES;
          const uint32_t ph = (it / SL::STAGES) & 1;
          mbar_wait(full_bar(s), ph);
          tc_fence_after();
          const uint32_t st = smem_base + s * SL::STAGE;
          const uint32_t a_hi = st, a_lo = st + SL::A_TILE, b_hi = st + 2 * SL::A_TILE;
          if (elect_one()) {
            const bool no_blo = skip_b1lo && (kb_lo + kq) < nk1;
#pragma unroll
            for (int j = 0; j < TBK / UMMA_K; ++j) {
              const uint64_t dah = make_smem_desc(a_hi + j * A_KSTEP, A_LBO, A_SBO, A_LT);
              const uint64_t dal = make_smem_desc(a_lo + j * A_KSTEP, A_LBO, A_SBO, A_LT);
              const uint64_t dbh = make_smem_desc(b_hi + j * B_KSTEP, B_LBO, B_SBO, B_LT);
              const uint32_t acc = (kq != c * KCr || j != 0) ? 1u : 0u;
              if (no_blo) {      // exactly-TF32 B (lo identically zero, its tile not loaded): main and one cross term
                umma_tf32(t_main, dah, dbh, idesc, acc);
                umma_tf32(t_cross, dal, dbh, idesc, acc);
              } else if (g.merge != 1) {
                const uint64_t dbl = make_smem_desc(b_hi + SL::B_TILE + j * B_KSTEP, B_LBO, B_SBO, B_LT);
                umma_tf32(t_cross, dal, dbh, idesc, acc);
                umma_tf32(t_cross, dah, dbl, idesc, 1);
                umma_tf32(t_main, dah, dbh, idesc, acc);
              } else {
                umma_tf32(t_main, dah, dbh, idesc2, acc);       // [main | A_hi x B_lo]
                umma_tf32(t_cross, dal, dbh, idesc, 1);         // += A_lo x B_hi
              }
            }
            umma_commit(empty_bar(s));
          }
          __syncwarp();
        }
        if (elect_one()) umma_commit(tfull_bar(buf));
        __syncwarp();
      }
    }
  } else if (warp >= 4) {
    // ================= drain + epilogue warps: warp -> TMEM lane quarter q, 64-column half h =================
    const int q = warp & 3;
    const int h = (warp - 4) >> 2;
    constexpr int HC = 64;
    float* stg = reinterpret_cast<float*>(smem_raw + (stg_base - smem_u32(smem_raw))) + (warp - 4) * 32 * STG_LD;
    TcParams p = p0;
    if (AMODE == 2 && S > 1) {      // partial tiles go, unscaled, to the split-K scratch [batch*S][M][N]
      p.C = p0.colsum; p.C_lo = nullptr; p.c_sz = (long long)p0.M * p0.N; p.c_sm = p0.N;
      p.scale = 1.f; p.bias = nullptr; p.mask = nullptr; p.add = nullptr;
    }
    p.colsum = nullptr;
    uint32_t ck = 0;
    for (long long t = blockIdx.x; t < ntiles; t += gridDim.x) {
      int m0, n0, z, ks, kb_lo, kb_hi;
      tile_coords(t, m0, n0, z, ks, kb_lo, kb_hi);
      const int nchunks = (kb_hi - kb_lo + KCr - 1) / KCr;
      float acc[HC];
#pragma unroll
      for (int i = 0; i < HC; ++i) acc[i] = 0.f;
      if (EPI == 0 && p.add != nullptr) {
        // the epilogue reads add[m][n] (the cotangent already in the slot) in four dependent 8-row batches: pull this warp's
        // 32 x 64 block into L2 now, while the tensor core still works on the tile (lane <-> row, one 128-byte line per group)
        const int row = m0 + q * 32 + lane;
        if (row < p.M) {
          const float* a = p.add + (long long)(z * S + ks) * p.add_sz + (long long)row * p.c_sm + n0 + h * HC;
#pragma unroll
          for (int cc = 0; cc < HC / 32; ++cc)
            if (h * HC + cc * 32 < NB && n0 + h * HC + cc * 32 < p.N) asm volatile("prefetch.global.L2 [%0];" ::"l"(a + cc * 32));
        }
      }
      if (EPI == 1) {
        // same for the BatchNorm-JVP epilogue's operands (xhat, mask, first-term sum, skip pair)
        const int row = m0 + q * 32 + lane;
        if (row < p.M) {
          const long long i0 = (long long)row * p.c_sm + n0 + h * HC, zi = (long long)z * p.c_sz + i0;
#pragma unroll
          for (int cc = 0; cc < HC / 32; ++cc) {
            if (h * HC + cc * 32 < NB && n0 + h * HC + cc * 32 < p.N) {
              asm volatile("prefetch.global.L2 [%0];" ::"l"(bn.xhat + i0 + cc * 32));
              if (bn.mask) asm volatile("prefetch.global.L2 [%0];" ::"l"(bn.mask + i0 + cc * 32));
              if (bn.pre) asm volatile("prefetch.global.L2 [%0];" ::"l"(bn.pre + zi + cc * 32));
              if (bn.skip_hi) {
                asm volatile("prefetch.global.L2 [%0];" ::"l"(bn.skip_hi + zi + cc * 32));
                asm volatile("prefetch.global.L2 [%0];" ::"l"(bn.skip_lo + zi + cc * 32));
              }
            }
          }
        }
      }
      for (int c = 0; c < nchunks; ++c, ++ck) {
        const uint32_t buf = ck & 1, cph = (ck >> 1) & 1;
        mbar_wait(tfull_bar(buf), cph);
        tc_fence_after();
        const uint32_t tl = tmem_base + ((uint32_t)(q * 32) << 16) + buf * (2 * TSTRIDE) + h * HC;
#pragma unroll
        for (int cc = 0; cc < HC / 32; ++cc) {
          if (h * HC + cc * 32 < NB) {
            float w[32];
            tmem_ld32(tl + (uint32_t)((g.merge == 2 ? TSTRIDE : NB) + cc * 32), w);            // cross terms
#pragma unroll
            for (int i = 0; i < 32; ++i) acc[cc * 32 + i] += w[i];
            tmem_ld32(tl + (uint32_t)(cc * 32), w);                 // main term
#pragma unroll
            for (int i = 0; i < 32; ++i) acc[cc * 32 + i] += w[i];
          }
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(tempty_bar(buf));
      }
      if (EPI == 2) {
        // fold mode: each 32-column group of the tile belongs to one probe (output slot probe*S + slice)
#pragma unroll
        for (int cc = 0; cc < 2; ++cc) {
          const int grp = h * 2 + cc;
          const int probe = z * G + grp / g.cpb;
          if (probe < p0.batch)
            tile_epilogue<1, 8>(p, reinterpret_cast<float(&)[32]>(acc[cc * 32]), stg, m0, (grp % g.cpb) * 32, probe * S + ks, q, 0,
                                lane);
        }
      } else if (h * HC < NB) {
        const int zs = z * S + ks;
        if (EPI == 1) bn_tile_epilogue<2, 8>(p, bn, acc, stg, m0, n0, zs, q, h, lane);
        else tile_epilogue<2, 8>(p, acc, stg, m0, n0, zs, q, h, lane);
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)TMEM_COLS));
  }
}

// out[z][m*c_sm + n] = scale * sum_s ws[(z*S + s)][m][n] + add_scale * add[z][m*c_sm + n]      (fixed order: deterministic)
__global__ void conv_splitk_reduce_kernel(const float* __restrict__ ws, int S, long long MN, int N, float* __restrict__ out,
                                          long long c_sz, long long c_sm, float scale, const float* __restrict__ add,
                                          long long add_sz, float add_scale, long long total) {
  for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total; idx += (long long)gridDim.x * blockDim.x) {
    const long long z = idx / MN, i = idx - z * MN;
    const long long m = i / N;
    const int n = (int)(i - m * N);
    const float* w = ws + z * S * MN + i;
    float acc = 0.f;
    for (int s = 0; s < S; ++s) acc += __ldg(w + (long long)s * MN);
    const long long o = m * c_sm + n;
    float v = scale * acc;
    if (add) v = fmaf(add_scale, __ldg(add + z * add_sz + o), v);
    out[z * c_sz + o] = v;
  }
}

// 4-D map over an NHWC image batch [images, H, W, C]; box = `rows` consecutive OUTPUT pixels (whole rows / images of the
// Ho x Wo = H/stride x W/stride output grid) x 32 channels, traversed with element stride `stride` along W and H
int make_image_map(CUtensorMap* map, const float* ptr, int64_t images, int H, int W, int C, int rows, bool kmajor, int stride) {
  EncodeTiledFn enc = tc_get_encode();
  if (!enc) { set_error("cuTensorMapEncodeTiled unavailable"); return LIP_ERR_UNSUPPORTED; }
  LIP_REQUIRE(((uintptr_t)ptr & 15) == 0 && C % 32 == 0, "conv_tc: image not 16-byte aligned / channels not a multiple of 32");
  const int Ho = H / stride, Wo = W / stride;
  const int bw = Wo < rows ? Wo : rows;
  const int bh = Ho < rows / bw ? Ho : rows / bw;
  const int bn = rows / (bw * bh);
  cuuint64_t dims[4] = {(cuuint64_t)C, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)images};
  cuuint64_t strides[3] = {(cuuint64_t)C * 4, (cuuint64_t)W * C * 4, (cuuint64_t)H * W * C * 4};
  cuuint32_t box[4] = {32, (cuuint32_t)(bw * stride), (cuuint32_t)(bh * stride), (cuuint32_t)bn};
  cuuint32_t es[4] = {1, (cuuint32_t)stride, (cuuint32_t)stride, 1};
  CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, (void*)ptr, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   kmajor ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled (image) failed (%d): images=%lld H=%d W=%d C=%d box=%d,%d,%d stride=%d", (int)r,
              (long long)images, H, W, C, bw, bh, bn, stride);
    return LIP_ERR_CUDA;
  }
  return LIP_OK;
}

// `rows` consecutive pixels starting at a multiple of `rows` form a box of whole rows / whole images
bool box_tiles(int H, int W, int rows) {
  if (W >= rows) return W % rows == 0;
  if (rows % W != 0) return false;
  const int hh = rows / W;
  if (H >= hh) return H % hh == 0;
  return hh % H == 0;
}

int num_sms() {
  static int n = 0;
  if (n == 0) {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess) n = 148;
  }
  return n;
}

template <int NB, int AMODE, int EPI>
int launch_conv(const CUtensorMap* maps, const TcParams& p, const ConvGeo& g, const ConvBnEpilogue& bn, int64_t ntiles,
                cudaStream_t st) {
  using SL = ConvSmem<NB>;
  auto kern = conv_tc_kernel<NB, AMODE, EPI>;
  static bool attr_set = false;
  if (!attr_set) {
    LIP_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, SL::BYTES));
    attr_set = true;
  }
  const int sms = num_sms();
  const unsigned grid = (unsigned)(ntiles < sms ? ntiles : sms);
  kern<<<grid, TC_THREADS, SL::BYTES, st>>>(maps[0], maps[1], maps[2], maps[3], maps[4], maps[5], maps[6], maps[7], p, g, bn);
  LIP_LAUNCH_CHECK();
  return LIP_OK;
}

template <int AMODE>
int dispatch_conv(int nb, const CUtensorMap* maps, const TcParams& p, const ConvGeo& g, const ConvBnEpilogue& bn, int64_t ntiles,
                  cudaStream_t st) {
  if (g.fold > 1) return launch_conv<128, AMODE, 2>(maps, p, g, bn, ntiles, st);
  if (AMODE == 1 && bn.on) {
    if (nb == 32) return launch_conv<32, 1, 1>(maps, p, g, bn, ntiles, st);
    if (nb == 64) return launch_conv<64, 1, 1>(maps, p, g, bn, ntiles, st);
    return launch_conv<128, 1, 1>(maps, p, g, bn, ntiles, st);
  }
  if (nb == 32) return launch_conv<32, AMODE, 0>(maps, p, g, bn, ntiles, st);
  if (nb == 64) return launch_conv<64, AMODE, 0>(maps, p, g, bn, ntiles, st);
  return launch_conv<128, AMODE, 0>(maps, p, g, bn, ntiles, st);
}

inline int tile_width(int64_t N) { return N <= 32 ? 32 : (N <= 64 ? 64 : 128); }
// probes folded into one 128-column tile when the A operand is shared by all probes (N = 32 -> 4, N = 64 -> 2)
inline int fold_of(int64_t N, int64_t batch) {
  static const int off = getenv("LIP_CONV_TC_FOLD") ? (atoi(getenv("LIP_CONV_TC_FOLD")) == 0) : 0;
  if (off || batch < 2 || (N != 32 && N != 64)) return 1;
  return (int)(128 / N);
}

int merge_default() {
  static const int m = getenv("LIP_CONV_TC_MERGE") ? atoi(getenv("LIP_CONV_TC_MERGE")) : 1;
  return m;
}

void fill_params(TcParams* p, int64_t M, int64_t N, int64_t K1, int64_t K2, int64_t batch) {
  memset(p, 0, sizeof(*p));
  p->M = (int)M; p->N = (int)N; p->K1 = (int)K1; p->K2 = (int)K2; p->batch = (int)batch;
  p->kc = getenv("LIP_TC_KC") ? atoi(getenv("LIP_TC_KC")) : KC;
  p->scale = 1.f;
}

}  // namespace

bool conv_tc_supported(int H, int W, int C, int N, int kh, int kw, int stride, int pad) {
  static const int off = getenv("LIP_CONV_TC") ? (atoi(getenv("LIP_CONV_TC")) == 0) : 0;
  static const int no_s2 = getenv("LIP_CONV_TC_S2") ? (atoi(getenv("LIP_CONV_TC_S2")) == 0) : 0;
  if (off || !tc_available()) return false;
  if ((stride != 1 && stride != 2) || kh != kw || (kh != 1 && kh != 3) || pad < 0 || pad > kh / 2) return false;
  if (stride == 2 && (no_s2 || (H % 2) || (W % 2))) return false;
  if (C % 32 != 0 || N % 32 != 0 || C > 512 || N > 512) return false;
  const int Ho = H / stride, Wo = W / stride;
  // 128- and 32-pixel runs of the output grid are boxes of whole rows / images; the transposed conv of a strided unit runs
  // on the zero-upsampled delta image, i.e. on the H x W grid
  return box_tiles(Ho, Wo, 128) && box_tiles(Ho, Wo, 32) && box_tiles(H, W, 128) && W <= 128 && H <= 128;
}

int conv_tc(const ConvTcProblem& c, cudaStream_t st) {
  const int sd = c.stride;
  LIP_REQUIRE((sd == 1 || sd == 2) && !(c.transposed && sd != 1), "conv_tc: stride must be 1 or 2 (1 for a transposed conv)");
  const int Ho = c.H / sd, Wo = c.W / sd;
  const int64_t P = (int64_t)Ho * Wo, R = c.imgs * P, Kc = (int64_t)c.kh * c.kw * c.C;
  LIP_REQUIRE(c.A1.hi && c.A1.lo && c.B1.hi && c.B1.lo && c.C_out, "conv_tc: null operand");
  LIP_REQUIRE(R < (1ll << 31) && c.batch * c.imgs < (1ll << 31), "conv_tc: problem too large for 32-bit tile indices");
  const bool dual = c.A2.hi != nullptr;
  CUtensorMap maps[8];
  const ConvTcImage* as[2] = {&c.A1, dual ? &c.A2 : &c.A1};
  const TcOperand* bs[2] = {&c.B1, dual ? &c.B2 : &c.B1};
  for (int i = 0; i < 2; ++i) {
    const int64_t images = as[i]->batched ? c.batch * c.imgs : c.imgs;
    int rc = make_image_map(&maps[4 * i], as[i]->hi, images, c.H, c.W, c.C, 128, true, sd);
    if (!rc) rc = make_image_map(&maps[4 * i + 1], as[i]->lo, images, c.H, c.W, c.C, 128, true, sd);
    const bool bb = (i == 0 ? c.b1_batched : c.b2_batched) != 0;
    if (!rc) rc = tc_make_map(&maps[4 * i + 2], bs[i]->hi, false, c.N, Kc, bs[i]->ld, bs[i]->sz, bb ? c.batch : 1, 32);
    if (!rc) rc = tc_make_map(&maps[4 * i + 3], bs[i]->lo, false, c.N, Kc, bs[i]->ld, bs[i]->sz, bb ? c.batch : 1, 32);
    if (rc) return rc;
  }
  TcParams p;
  fill_params(&p, R, c.N, Kc, dual ? Kc : 0, c.batch);
  p.a1_batched = c.A1.batched; p.b1_batched = c.b1_batched;
  p.a2_batched = dual ? c.A2.batched : 0; p.b2_batched = dual ? c.b2_batched : 0;
  p.C = c.C_out; p.C_lo = c.C_lo; p.c_sz = c.c_sz; p.c_sm = c.c_sm;
  p.scale = c.epi.scale;
  p.bias = c.epi.bias; p.bias_sz = c.epi.bias_sz;
  p.mask = c.epi.mask; p.mask_sm = c.epi.mask_sm;
  p.add = c.epi.add; p.add_sz = c.epi.add_sz; p.add_scale = c.epi.add_scale;
  p.b1_lo_nz = c.B1.lo_nz;
  ConvGeo g{};
  g.W = Wo; g.H = Ho; g.P = (int)P; g.stride = sd; g.C = c.C; g.cblocks = c.C / 32; g.kw = c.kw;
  g.sgn = c.transposed ? -1 : 1;
  g.off_h = c.transposed ? c.pad : -c.pad; g.off_w = g.off_h;
  g.imgs = (int)c.imgs; g.Kc = (int)Kc; g.ksplit = 1; g.kb_per = 0;
  g.fold = 1; g.cpb = (int)(c.N / 32); g.merge = merge_default();
  if (c.fold_probes) {
    LIP_REQUIRE(!dual && !c.A1.batched && c.b1_batched && !c.bn.on && !c.C_lo, "conv_tc: probe folding needs one shared image");
    g.fold = fold_of(c.N, c.batch);
  }
  const int nb = g.fold > 1 ? 128 : tile_width(c.N);
  const int64_t ntiles = g.fold > 1 ? ceil_div(R, TBM) * ceil_div(c.batch, g.fold) : ceil_div(R, TBM) * ceil_div(c.N, nb) * c.batch;
  if (c.bn.on) {
    LIP_REQUIRE(c.C_lo && c.c_sm == c.N && c.bn.g && c.bn.xhat && c.bn.dscale && c.bn.dbeta && (!c.bn.skip_hi || c.bn.skip_lo),
                "conv_tc: the fused BatchNorm epilogue needs a (hi, lo) output with dense rows and all of g / xhat / dscale / dbeta");
  }
  return dispatch_conv<1>(nb, maps, p, g, c.bn, ntiles, st);
}

int64_t conv_wgrad_tc_splits(int64_t imgs, int Ho, int Wo, int C, int N, int kh, int kw, int64_t batch) {
  const int64_t nk = ceil_div(imgs * (int64_t)Ho * Wo, TBK);
  const int fold = fold_of(N, batch);
  const int64_t tiles = fold > 1 ? ceil_div((int64_t)kh * kw * C, TBM) * ceil_div(batch, fold)
                                 : ceil_div((int64_t)kh * kw * C, TBM) * ceil_div(N, tile_width(N)) * batch;
  int64_t S = ceil_div(2 * (int64_t)num_sms(), tiles);
  const int64_t smax = nk / 16 > 1 ? nk / 16 : 1;     // at least 16 k-blocks (two TMEM chunks) per slice
  if (S > smax) S = smax;
  if (S > 64) S = 64;
  if (S < 1) S = 1;
  const int64_t per = ceil_div(nk, S);
  return ceil_div(nk, per);
}

int conv_wgrad_tc(const ConvWgradTcProblem& c, cudaStream_t st) {
  const int sd = c.stride;
  LIP_REQUIRE(sd == 1 || sd == 2, "conv_wgrad_tc: stride must be 1 or 2");
  const int Ho = c.H / sd, Wo = c.W / sd;
  const int64_t P = (int64_t)Ho * Wo, R = c.imgs * P, Kc = (int64_t)c.kh * c.kw * c.C;
  LIP_REQUIRE(c.X_hi && c.X_lo && c.D.hi && c.D.lo && c.C_out, "conv_wgrad_tc: null operand");
  LIP_REQUIRE(R < (1ll << 31), "conv_wgrad_tc: problem too large for 32-bit indices");
  CUtensorMap maps[8];
  int rc = make_image_map(&maps[0], c.X_hi, c.imgs, c.H, c.W, c.C, 32, false, sd);
  if (!rc) rc = make_image_map(&maps[1], c.X_lo, c.imgs, c.H, c.W, c.C, 32, false, sd);
  if (!rc) rc = tc_make_map(&maps[2], c.D.hi, false, c.N, R, c.D.ld, c.D.sz, c.batch, 32);
  if (!rc) rc = tc_make_map(&maps[3], c.D.lo, false, c.N, R, c.D.ld, c.D.sz, c.batch, 32);
  if (rc) return rc;
  for (int i = 4; i < 8; ++i) maps[i] = maps[i - 4];
  const int64_t nk = ceil_div(R, TBK);
  int64_t S = conv_wgrad_tc_splits(c.imgs, Ho, Wo, c.C, c.N, c.kh, c.kw, c.batch);
  if (S > 1 && (c.ws == nullptr || c.ws_elems < c.batch * S * Kc * c.N)) {
    S = c.ws ? c.ws_elems / (c.batch * Kc * c.N) : 1;
    if (S < 1) S = 1;
  }
  const int64_t per = ceil_div(nk, S);
  S = ceil_div(nk, per);
  TcParams p;
  fill_params(&p, Kc, c.N, R, 0, c.batch);
  p.a1_batched = 0; p.b1_batched = 1;
  p.C = c.C_out; p.c_sz = c.c_sz; p.c_sm = c.c_sm;
  p.scale = c.epi.scale;
  p.add = c.epi.add; p.add_sz = c.epi.add_sz; p.add_scale = c.epi.add_scale;
  p.colsum = S > 1 ? c.ws : nullptr;          // (re-used field) split-K scratch
  ConvGeo g{};
  g.W = Wo; g.H = Ho; g.P = (int)P; g.stride = sd; g.C = c.C; g.cblocks = c.C / 32; g.kw = c.kw;
  g.sgn = 1; g.off_h = -c.pad; g.off_w = -c.pad;
  g.imgs = (int)c.imgs; g.Kc = (int)Kc; g.ksplit = (int)S; g.kb_per = (int)per;
  g.fold = fold_of(c.N, c.batch); g.cpb = (int)(c.N / 32); g.merge = merge_default();
  const int nb = g.fold > 1 ? 128 : tile_width(c.N);
  const int64_t ntiles = (g.fold > 1 ? ceil_div(Kc, TBM) * ceil_div(c.batch, g.fold) : ceil_div(Kc, TBM) * ceil_div(c.N, nb) * c.batch) * S;
  rc = dispatch_conv<2>(nb, maps, p, g, ConvBnEpilogue(), ntiles, st);
  if (rc || S == 1) return rc;
  const long long MN = Kc * c.N, total = MN * c.batch;
  long long grid = (total + 255) / 256;
  if (grid > 148ll * 16) grid = 148ll * 16;
  conv_splitk_reduce_kernel<<<(unsigned)grid, 256, 0, st>>>(c.ws, (int)S, MN, (int)c.N, c.C_out, c.c_sz, c.c_sm, c.epi.scale,
                                                            c.epi.add, c.epi.add_sz, c.epi.add_scale, total);
  LIP_LAUNCH_CHECK();
  return LIP_OK;
}

namespace {
// up[(img, 2y, 2x), c] = d[(img, y, x), c] for both TF32 parts (`up` is zero-filled beforehand)
__global__ void upsample2_kernel(const float* __restrict__ dh, const float* __restrict__ dl, float* __restrict__ uh,
                                 float* __restrict__ ul, long long total, int Ho, int Wo, int C) {
  for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total; idx += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(idx % C);
    long long t = idx / C;
    const int x = (int)(t % Wo); t /= Wo;
    const int y = (int)(t % Ho);
    const long long img = t / Ho;
    const long long o = ((img * (2 * Ho) + 2 * y) * (2 * Wo) + 2 * x) * C + c;
    uh[o] = dh[idx];
    ul[o] = dl[idx];
  }
}
}  // namespace

int conv_tc_upsample2(const float* d_hi, const float* d_lo, float* up_hi, float* up_lo, int64_t images, int Ho, int Wo, int C,
                      cudaStream_t st) {
  const long long total = images * (long long)Ho * Wo * C;
  LIP_CHECK_CUDA(cudaMemsetAsync(up_hi, 0, sizeof(float) * (size_t)total * 4, st));
  LIP_CHECK_CUDA(cudaMemsetAsync(up_lo, 0, sizeof(float) * (size_t)total * 4, st));
  long long grid = (total + 255) / 256;
  if (grid > 148ll * 32) grid = 148ll * 32;
  upsample2_kernel<<<(unsigned)grid, 256, 0, st>>>(d_hi, d_lo, up_hi, up_lo, total, Ho, Wo, C);
  LIP_LAUNCH_CHECK();
  return LIP_OK;
}

}  // namespace lip

// ---- self test / microbenchmark: tcgen05 implicit-GEMM conv vs the fp32 SIMT implicit GEMM on random data -------------------
namespace {
__global__ void conv_fill_random_kernel(float* x, long long n, unsigned seed, float scale) {
  long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (i >= n) return;
  unsigned s = (unsigned)(i * 2654435761u) ^ seed;
  s ^= s >> 16; s *= 0x7feb352du; s ^= s >> 15; s *= 0x846ca68bu; s ^= s >> 16;
  x[i] = scale * ((float)(s & 0xFFFFFF) / 8388608.f - 1.f);
}
__global__ void conv_rel_err_kernel(const float* a, const float* b, long long n, float* num, float* den) {
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

// role 0: JVP (dual K: shared image x per-probe kernels + per-probe image x shared kernel), 1: kernel gradient,
// 2: delta back-propagation (transposed conv).  Returns the relative L2 error against the SIMT implicit GEMM and, with
// iters > 0, the mean time of one tensor-core call (ms_tc) and of one SIMT call (ms_simt).
extern "C" int lip_selftest_conv_tc(int32_t role, int64_t imgs, int32_t H, int32_t W, int32_t cin, int32_t cout, int32_t ksz,
                                    int32_t stride, int64_t batch, int32_t iters, float* rel_err, float* ms_tc, float* ms_simt,
                                    lip_stream_t stream) {
  using namespace lip;
  LIP_REQUIRE(role >= 0 && role <= 3 && imgs > 0 && batch > 0 && rel_err && (stride == 1 || stride == 2),
              "conv selftest: bad argument");
  const int pad = stride == 1 ? (ksz - 1) / 2 : 0;      // XLA 'SAME' (smaller half first)
  if (!conv_tc_supported(H, W, cin, cout, ksz, ksz, stride, pad)) {
    set_error("conv selftest: shape not supported by the tcgen05 conv path");
    return LIP_ERR_UNSUPPORTED;
  }
  cudaStream_t st = (cudaStream_t)stream;
  const int Ho = H / stride, Wo = W / stride;
  const int64_t Pi = (int64_t)H * W, P = (int64_t)Ho * Wo, R = imgs * P, Ri = imgs * Pi;
  const int64_t Kc = (int64_t)ksz * ksz * cin, Kt = (int64_t)ksz * ksz * cout;
  // buffers: X [imgs,H,W,cin] shared image, T [batch,imgs,H,W,cin] per-probe image, Wk [Kc,cout] shared kernel,
  // dW [batch,Kc,cout] per-probe kernels, Dh [batch,R,cout] deltas, Wt [Kt,cin] re-laid kernel
  const int64_t nX = Ri * cin, nT = batch * Ri * cin, nW = Kc * cout, ndW = batch * Kc * cout, nD = batch * R * cout, nWt = Kt * cin;
  const int64_t nOut = (role == 0 || role == 3) ? nD : (role == 1 ? ndW : nT);
  const int64_t nUp = stride == 2 ? batch * Ri * cout : 1;
  const int64_t S = conv_wgrad_tc_splits(imgs, Ho, Wo, cin, cout, ksz, ksz, batch);
  const int64_t nws = batch * S * Kc * cout;
  std::vector<float*> bufs;
  auto alloc = [&](int64_t n) -> float* {
    float* q = nullptr;
    if (cudaMalloc(&q, sizeof(float) * (size_t)(n + 64)) != cudaSuccess) return nullptr;
    bufs.push_back(q);
    return q;
  };
  auto free_all = [&]() { for (float* q : bufs) cudaFree(q); };
  float *X = alloc(nX), *Xh = alloc(nX), *Xl = alloc(nX), *T = alloc(nT), *Th = alloc(nT), *Tl = alloc(nT);
  float *Wk = alloc(nW), *Wh = alloc(nW), *Wl = alloc(nW), *dW = alloc(ndW), *dWh = alloc(ndW), *dWl = alloc(ndW);
  float *Dh = alloc(nD), *Dhh = alloc(nD), *Dhl = alloc(nD), *Wt = alloc(nWt), *Wth = alloc(nWt), *Wtl = alloc(nWt);
  float *C0 = alloc(nOut), *C1 = alloc(nOut), *addv = alloc(nOut), *ws = alloc(nws), *stats = alloc(2);
  float *Uh = alloc(nUp), *Ul = alloc(nUp);
  for (float* q : {X, Xh, Xl, T, Th, Tl, Wk, Wh, Wl, dW, dWh, dWl, Dh, Dhh, Dhl, Wt, Wth, Wtl, C0, C1, addv, ws, stats, Uh, Ul}) {
    if (!q) { free_all(); set_error("conv selftest: out of device memory"); return LIP_ERR_CUDA; }
  }
  auto fill = [&](float* q, int64_t n, unsigned seed) {
    conv_fill_random_kernel<<<(unsigned)ceil_div(n, 256), 256, 0, st>>>(q, n, seed, 1.f);
  };
  fill(X, nX, 3u); fill(T, nT, 5u); fill(Wk, nW, 7u); fill(dW, ndW, 11u); fill(Dh, nD, 13u); fill(Wt, nWt, 17u); fill(addv, nOut, 19u);
  int rc = tf32_split(X, nX, Xh, Xl, nX, 1, nX, st);
  if (!rc) rc = tf32_split(T, nT, Th, Tl, nT, 1, nT, st);
  if (!rc) rc = tf32_split(Wk, nW, Wh, Wl, nW, 1, nW, st);
  if (!rc) rc = tf32_split(dW, ndW, dWh, dWl, ndW, 1, ndW, st);
  if (!rc) rc = tf32_split(Dh, nD, Dhh, Dhl, nD, 1, nD, st);
  if (!rc) rc = tf32_split(Wt, nWt, Wth, Wtl, nWt, 1, nWt, st);
  ConvGather cg;
  cg.Hi = H; cg.Wi = W; cg.pad_h = pad; cg.pad_w = pad; cg.stride = stride; cg.kh = ksz; cg.kw = ksz; cg.Ho = Ho; cg.Wo = Wo;
  GemmProblem sp;
  ConvTcProblem tp;
  ConvWgradTcProblem wp;
  if (role == 0 || role == 3) {
    cg.mode = 1; cg.C = cin;
    sp.M = R; sp.N = cout; sp.K = Kc; sp.batch = batch;
    sp.A1.ptr = X; sp.A1.sz = 0; sp.A1.conv = cg;
    sp.B1 = {dW, Kc * cout, cout, 1};
    sp.A2.ptr = T; sp.A2.sz = Ri * cin; sp.A2.conv = cg;
    sp.B2 = {Wk, 0, cout, 1};
    sp.K2 = Kc;
    sp.C = C0; sp.c_sz = R * cout; sp.c_sm = cout;
    tp.imgs = imgs; tp.H = H; tp.W = W; tp.C = cin; tp.N = cout; tp.kh = ksz; tp.kw = ksz; tp.pad = pad; tp.batch = batch;
    tp.stride = stride;
    tp.A1 = {Xh, Xl, 0};
    tp.B1.hi = dWh; tp.B1.lo = dWl; tp.B1.sz = Kc * cout; tp.B1.ld = cout; tp.B1.major_k = 0; tp.b1_batched = 1;
    tp.A2 = {Th, Tl, 1};
    tp.B2.hi = Wh; tp.B2.lo = Wl; tp.B2.sz = 0; tp.B2.ld = cout; tp.B2.major_k = 0; tp.b2_batched = 0;
    tp.C_out = C1; tp.c_sz = R * cout; tp.c_sm = cout;
    if (role == 3) {          // first JVP term alone, probes folded into the tile width
      sp.A2 = GemmOperand(); sp.B2 = GemmOperand(); sp.K2 = 0;
      tp.A2 = ConvTcImage(); tp.B2 = TcOperand();
      tp.fold_probes = 1;
    }
  } else if (role == 1) {
    cg.mode = 2; cg.C = cin;
    sp.M = Kc; sp.N = cout; sp.K = R; sp.batch = batch;
    sp.A1.ptr = X; sp.A1.sz = 0; sp.A1.conv = cg;
    sp.B1 = {Dh, R * cout, cout, 1};
    sp.C = C0; sp.c_sz = Kc * cout; sp.c_sm = cout;
    sp.epi.scale = 0.5f; sp.epi.add = addv; sp.epi.add_sz = Kc * cout; sp.epi.add_scale = 0.25f;
    wp.imgs = imgs; wp.H = H; wp.W = W; wp.C = cin; wp.N = cout; wp.kh = ksz; wp.kw = ksz; wp.pad = pad; wp.batch = batch;
    wp.stride = stride;
    wp.X_hi = Xh; wp.X_lo = Xl;
    wp.D.hi = Dhh; wp.D.lo = Dhl; wp.D.sz = R * cout; wp.D.ld = cout; wp.D.major_k = 0;
    wp.C_out = C1; wp.c_sz = Kc * cout; wp.c_sm = cout;
    wp.epi = sp.epi;
    wp.ws = ws; wp.ws_elems = nws;
  } else {
    cg.mode = 3; cg.C = cout;
    sp.M = Ri; sp.N = cin; sp.K = Kt; sp.batch = batch;
    sp.A1.ptr = Dh; sp.A1.sz = R * cout; sp.A1.conv = cg;
    sp.B1 = {Wt, 0, cin, 1};
    sp.C = C0; sp.c_sz = Ri * cin; sp.c_sm = cin;
    sp.epi.add = addv; sp.epi.add_sz = Ri * cin; sp.epi.add_scale = 1.f;
    // a strided unit back-propagates through the zero-upsampled delta image (stride-1 transposed conv on the H x W grid)
    tp.imgs = imgs; tp.H = H; tp.W = W; tp.C = cout; tp.N = cin; tp.kh = ksz; tp.kw = ksz; tp.pad = pad; tp.batch = batch;
    tp.transposed = 1;
    if (stride == 2) tp.A1 = {Uh, Ul, 1}; else tp.A1 = {Dhh, Dhl, 1};
    tp.B1.hi = Wth; tp.B1.lo = Wtl; tp.B1.sz = 0; tp.B1.ld = cin; tp.B1.major_k = 0; tp.b1_batched = 0;
    tp.C_out = C1; tp.c_sz = Ri * cin; tp.c_sm = cin;
    if (getenv("LIP_SELFTEST_NOADD")) { sp.epi.add = nullptr; sp.epi.add_scale = 0.f; }
    tp.epi = sp.epi;
  }
  sp.epi.C_lo = nullptr;
  cg.finalize();
  sp.A1.conv.finalize();
  if (sp.A2.ptr) sp.A2.conv.finalize();
  auto run_tc = [&]() {
    if (role == 1) return conv_wgrad_tc(wp, st);
    if (role == 2 && stride == 2) {
      int r2 = conv_tc_upsample2(Dhh, Dhl, Uh, Ul, batch * imgs, Ho, Wo, cout, st);
      if (r2) return r2;
    }
    return conv_tc(tp, st);
  };
  if (!rc) rc = gemm_simt(sp, st);
  if (!rc) rc = run_tc();
  float h[2] = {0.f, 0.f};
  if (!rc) {
    cudaMemsetAsync(stats, 0, 8, st);
    conv_rel_err_kernel<<<256, 256, 0, st>>>(C1, C0, nOut, stats, stats + 1);
    cudaError_t e = cudaMemcpyAsync(h, stats, 8, cudaMemcpyDeviceToHost, st);
    if (e == cudaSuccess) e = cudaStreamSynchronize(st);
    if (e != cudaSuccess) { set_error("conv selftest: %s", cudaGetErrorString(e)); rc = LIP_ERR_CUDA; }
  }
  if (!rc && iters > 0 && ms_tc && ms_simt) {
    cudaEvent_t e0, e1, e2;
    cudaEventCreate(&e0); cudaEventCreate(&e1); cudaEventCreate(&e2);
    cudaEventRecord(e0, st);
    for (int i = 0; i < iters && !rc; ++i) rc = run_tc();
    cudaEventRecord(e1, st);
    for (int i = 0; i < iters && !rc; ++i) rc = gemm_simt(sp, st);
    cudaEventRecord(e2, st);
    cudaError_t e = cudaEventSynchronize(e2);
    if (e != cudaSuccess) { set_error("conv selftest (timing): %s", cudaGetErrorString(e)); rc = LIP_ERR_CUDA; }
    float a = 0.f, b = 0.f;
    cudaEventElapsedTime(&a, e0, e1); cudaEventElapsedTime(&b, e1, e2);
    *ms_tc = a / iters; *ms_simt = b / iters;
    cudaEventDestroy(e0); cudaEventDestroy(e1); cudaEventDestroy(e2);
  }
  free_all();
  if (rc) return rc;
  *rel_err = h[1] > 0.f ? sqrtf(h[0] / h[1]) : (h[0] > 0.f ? 1e30f : 0.f);
  return LIP_OK;
}
