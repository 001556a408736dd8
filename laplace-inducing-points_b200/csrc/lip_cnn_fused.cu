// lip_cnn_fused.cu — the conv stages of a small-channel conv stage program (LeNet5, src/scalemodels.py:11-49) without a patch buffer.
//
// lip_cnn.cu runs a conv stage as im2col + GEMM; for LeNet5's 1 -> 6 and 6 -> 16 channel 5x5 convs that costs a 3 GB patch buffer of
// the tangent image and a 3 GB `delta . W^T` column buffer per 256-probe call, and GEMMs whose N = 6 / 16 wastes the tiles
// (profiles/r02_launches_lenet5_summary.txt: 18 of 20.4 ms).  Here one (probe, image) pair lives in shared memory and the whole
// stage is one kernel per direction:
//   conv5_jvp_pool_kernel   T_out[b, m] = avgpool2( phi'(h) * ( conv(X[m], dW[b]) + conv(T_in[b, m], W) + db[b] ) )
//                           (src/ggn.py:133-144 JVP through nn.Conv + relu + avg_pool of src/scalemodels.py:24-36)
//   conv5_vjp_kernel        d = phi'(h) * unpool(t_in[b, m]) / 4;   gW[b] += patches(X[m])^T d;   gb[b] += colsum d;
//                           g_in[b, m] = transposed conv of d with W            (the VJP of the same ops)
// Geometry is a template parameter (all index arithmetic is compile-time); images are staged channel-planar and zero-padded so the
// inner loops read whole rows with 8 / 16-byte shared-memory loads and keep a register tile of (2x2 pool window) x channels (JVP),
// (5 taps) x (4 - 6 channels) (kernel gradient) or (7 pixels) x (input channels) (delta back-propagation).  Kernel / bias gradients
// are accumulated in registers over the CTA's images, reduced in a fixed order through shared memory and written as per-CTA
// partials; cnn_part_finish_kernel sums the partials in a fixed order and applies scale / +alpha V: deterministic, no atomics.
// Every product is an exact fp32 FMA (SIMT): the channel counts 1 / 6 / 16 are far below any tensor-core tile.
#include <stdlib.h>

#include "lip_model.cuh"

namespace lip {
namespace {

template <int CIN_, int COUT_, int HI_, int WI_, int PAD_>
struct Geo {
  static constexpr int CIN = CIN_, COUT = COUT_, HI = HI_, WI = WI_, PAD = PAD_, KS = 5;
  static constexpr int HO = HI + 2 * PAD - KS + 1, WO = WI + 2 * PAD - KS + 1, HP = HO / 2, WP = WO / 2;
  // zero-padded planar input image in shared memory; the row stride is even (8-byte row loads) and not a multiple of 32 floats
  // (threads that read different rows of one plane would otherwise all hit the same banks)
  static constexpr int HS = HI + 2 * PAD, WS0 = (WI + 2 * PAD + 1) & ~1, WS = (WS0 % 32 == 0) ? WS0 + 2 : WS0;
  static constexpr int PLANE = HS * WS;
  static constexpr int KK = KS * KS * CIN;                                  // rows of the flax HWIO kernel [(dy, dx, ci), co]
  static constexpr int CP = (COUT + 3) & ~3;                                // channel count padded to whole float4
  static_assert(HO % 2 == 0 && WO % 2 == 0, "2x2 average pool needs an even conv output");
  static bool match(const ConvStage& s) {
    return s.type == 1 && s.pool == 1 && s.kh == KS && s.kw == KS && s.cin == CIN && s.cout == COUT && s.Hi == HI && s.Wi == WI &&
           s.pad == PAD;
  }
};
using GeoC1 = Geo<1, 6, 28, 28, 2>;      // LeNet5 conv1: 28x28x1 -> 28x28x6 -> pool 14x14x6
using GeoC2 = Geo<6, 16, 14, 14, 0>;     // LeNet5 conv2: 14x14x6 -> 10x10x16 -> pool 5x5x16

template <int N>
__device__ __forceinline__ void lds_row(const float* p, float (&r)[N]) {      // p 8-byte aligned, N even
  static_assert(N % 2 == 0, "row length must be even");
#pragma unroll
  for (int i = 0; i < N; i += 2) {
    const float2 v = *reinterpret_cast<const float2*>(p + i);
    r[i] = v.x; r[i + 1] = v.y;
  }
}
template <int N>
__device__ __forceinline__ void lds_vec(const float* p, float (&r)[N]) {      // p 16-byte aligned, N in {4, 6, 8}
  static_assert(N == 4 || N == 6 || N == 8, "unsupported vector width");
  const float4 v = *reinterpret_cast<const float4*>(p);
  r[0] = v.x; r[1] = v.y; r[2] = v.z; r[3] = v.w;
  if constexpr (N == 6) {
    const float2 u = *reinterpret_cast<const float2*>(p + 4);
    r[4] = u.x; r[5] = u.y;
  }
  if constexpr (N == 8) {
    const float4 u = *reinterpret_cast<const float4*>(p + 4);
    r[4] = u.x; r[5] = u.y; r[6] = u.z; r[7] = u.w;
  }
}

template <int N>
__device__ __forceinline__ void ldg_vec(const float* p, float (&r)[N]) {      // N = 4: p 16-byte aligned; N = 6: p 8-byte aligned
  static_assert(N == 4 || N == 6, "unsupported vector width");
  if constexpr (N == 4) {
    const float4 v = __ldg(reinterpret_cast<const float4*>(p));
    r[0] = v.x; r[1] = v.y; r[2] = v.z; r[3] = v.w;
  } else {
#pragma unroll
    for (int i = 0; i < 6; i += 2) {
      const float2 v = __ldg(reinterpret_cast<const float2*>(p + i));
      r[i] = v.x; r[i + 1] = v.y;
    }
  }
}

// ---- JVP: conv + bias tangent + activation mask + 2x2 average pool ---------------------------------------------------------------
struct ConvJvpArgs {
  const float* X;       // [M, HI, WI, CIN]      stage input at the bound points
  const float* T;       // [B, M, HI, WI, CIN]   tangent of the stage input (DUAL)
  const float* W;       // [KK, COUT]            bound kernel (DUAL)
  const float* V;       // probe block at the stage's kernel offset (probe stride ldv)
  const float* Vb;      // ... at the bias offset
  const float* dphi;    // [M, HO, WO, COUT]
  float* out;           // [B, M, HP, WP, COUT]
  long long ldv;
  int M, IMG, ROUNDS;   // images staged per round, rounds per CTA (the kernels stay in shared memory across rounds)
};

template <class G, bool DUAL>
constexpr int jvp_smem_floats(int IMG) {
  return G::KK * G::CP * (DUAL ? 2 : 1) + 16 + IMG * G::CIN * G::PLANE * (DUAL ? 2 : 1);
}

template <class G, int CO_T, bool DUAL, int NT>
__global__ void __launch_bounds__(NT, 2) conv5_jvp_pool_kernel(ConvJvpArgs a) {
  constexpr int CIN = G::CIN, COUT = G::COUT, KS = G::KS, CP = G::CP, KK = G::KK, PLANE = G::PLANE, WS = G::WS;
  constexpr int NCG = COUT / CO_T, NW = G::HP * G::WP;
  static_assert(COUT % CO_T == 0 && (CO_T == 4 || CO_T == 6 || CO_T == 8), "channel group");
  extern __shared__ __align__(16) float sm[];
  float* sdW = sm;                                   // [KK][CP]  probe's kernel tangent
  float* sW = sdW + KK * CP;                         // [KK][CP]  bound kernel (DUAL)
  float* sdb = sW + (DUAL ? KK * CP : 0);            // [16]
  float* sX = sdb + 16;                              // [IMG][CIN][HS][WS]
  float* sT = sX + a.IMG * CIN * PLANE;              // same (DUAL)
  const int tid = threadIdx.x;
  const long long b = blockIdx.y;

  for (int i = tid; i < a.IMG * CIN * PLANE * (DUAL ? 2 : 1); i += NT) sX[i] = 0.f;
  {
    const float* dW = a.V + b * a.ldv;
#pragma unroll 8
    for (int i = tid; i < KK * CP; i += NT) {
      const int k = i / CP, c = i - k * CP;
      sdW[i] = c < COUT ? __ldg(dW + k * COUT + c) : 0.f;
      if (DUAL) sW[i] = c < COUT ? __ldg(a.W + k * COUT + c) : 0.f;
    }
    if (tid < 16) sdb[tid] = tid < COUT ? __ldg(a.Vb + b * a.ldv + tid) : 0.f;
  }
  __syncthreads();

  constexpr int PER = G::HI * G::WI * CIN;
  static_assert(PER % 4 == 0, "images are staged with 16-byte loads");
  for (int rd = 0; rd < a.ROUNDS; ++rd) {
  const int m0 = (blockIdx.x * a.ROUNDS + rd) * a.IMG;
  const int nimg = a.M - m0 < a.IMG ? a.M - m0 : a.IMG;
  if (nimg <= 0) break;
  if (rd > 0) __syncthreads();
  {
    const float4* Xg = reinterpret_cast<const float4*>(a.X + (long long)m0 * PER);
    const float4* Tg = DUAL ? reinterpret_cast<const float4*>(a.T + (b * a.M + m0) * PER) : nullptr;
    const int total4 = nimg * (PER / 4);
    constexpr int U = DUAL ? 2 : 4;
    for (int i0 = tid; i0 < total4; i0 += NT * U) {
      float4 vx[U], vt[U];
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const int i = i0 + u * NT;
        if (i < total4) {
          vx[u] = __ldg(Xg + i);
          if (DUAL) vt[u] = __ldg(Tg + i);
        }
      }
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const int i = i0 + u * NT;
        if (i < total4) {
          const float ex[4] = {vx[u].x, vx[u].y, vx[u].z, vx[u].w};
          const float et[4] = {DUAL ? vt[u].x : 0.f, DUAL ? vt[u].y : 0.f, DUAL ? vt[u].z : 0.f, DUAL ? vt[u].w : 0.f};
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const int e = 4 * i + j;
            const int img = e / PER, r = e - img * PER;
            const int pix = r / CIN, ci = r - pix * CIN;
            const int y = pix / G::WI, x = pix - y * G::WI;
            const int o = (img * CIN + ci) * PLANE + (y + G::PAD) * WS + x + G::PAD;
            sX[o] = ex[j];
            if (DUAL) sT[o] = et[j];
          }
        }
      }
    }
  }
  __syncthreads();

  // one item = (image pair, pool window, channel group): a register tile of 2 images x (2x2 positions) x CO_T channels, so that
  // every kernel vector read from shared memory feeds 8 * CO_T FMAs (the shared-memory pipe delivers 128 B/clk/SM: with 4
  // positions per thread the kernel reads alone would bound the loop).  The two operand pairs (X, dW) and (T, W) run one after
  // the other over the same accumulators, which halves the row registers.
  const int npair = (nimg + 1) >> 1;
  const int total = npair * NW * NCG;
  for (int it = tid; it < total; it += NT) {
    const int cg = it % NCG, w = it / NCG;
    const int pr = w / NW, wi = w - pr * NW;
    const int yp = wi / G::WP, xp = wi - yp * G::WP;
    const int co0 = cg * CO_T;
    float acc[2][2][2][CO_T];
#pragma unroll
    for (int g = 0; g < 2; ++g)
#pragma unroll
      for (int i = 0; i < 2; ++i)
#pragma unroll
        for (int j = 0; j < 2; ++j)
#pragma unroll
          for (int c = 0; c < CO_T; ++c) acc[g][i][j][c] = 0.f;
    const int woff = 2 * pr * CIN * PLANE + 2 * yp * WS + 2 * xp;
#pragma unroll 1
    for (int pass = 0; pass < (DUAL ? 2 : 1); ++pass) {
      const float* src = (pass ? sT : sX) + woff;
      const float* wsm = (pass ? sW : sdW) + co0;
#pragma unroll 1
      for (int ci = 0; ci < CIN; ++ci) {
        const float* p0 = src + ci * PLANE;
        const float* p1 = p0 + CIN * PLANE;
        float ra0[KS + 1], rb0[KS + 1], ra1[KS + 1], rb1[KS + 1];
        lds_row(p0, ra0);
        lds_row(p1, ra1);
#pragma unroll
        for (int dy = 0; dy < KS; ++dy) {
          lds_row(p0 + (dy + 1) * WS, rb0);
          lds_row(p1 + (dy + 1) * WS, rb1);
#pragma unroll
          for (int dx = 0; dx < KS; ++dx) {
            float wv[CO_T];
            lds_vec(wsm + ((dy * KS + dx) * CIN + ci) * CP, wv);
#pragma unroll
            for (int c = 0; c < CO_T; ++c) {
              acc[0][0][0][c] = fmaf(ra0[dx], wv[c], acc[0][0][0][c]);
              acc[0][0][1][c] = fmaf(ra0[dx + 1], wv[c], acc[0][0][1][c]);
              acc[0][1][0][c] = fmaf(rb0[dx], wv[c], acc[0][1][0][c]);
              acc[0][1][1][c] = fmaf(rb0[dx + 1], wv[c], acc[0][1][1][c]);
              acc[1][0][0][c] = fmaf(ra1[dx], wv[c], acc[1][0][0][c]);
              acc[1][0][1][c] = fmaf(ra1[dx + 1], wv[c], acc[1][0][1][c]);
              acc[1][1][0][c] = fmaf(rb1[dx], wv[c], acc[1][1][0][c]);
              acc[1][1][1][c] = fmaf(rb1[dx + 1], wv[c], acc[1][1][1][c]);
            }
          }
#pragma unroll
          for (int j = 0; j <= KS; ++j) { ra0[j] = rb0[j]; ra1[j] = rb1[j]; }
        }
      }
    }
    // bias tangent, activation mask, 2x2 mean
#pragma unroll
    for (int g = 0; g < 2; ++g) {
      if (2 * pr + g >= nimg) break;          // odd tail: the second image of the last pair is not there
      const long long m = m0 + 2 * pr + g;
      float s[CO_T];
#pragma unroll
      for (int c = 0; c < CO_T; ++c) s[c] = 0.f;
#pragma unroll
      for (int oy = 0; oy < 2; ++oy)
#pragma unroll
        for (int ox = 0; ox < 2; ++ox) {
          const float* dp = a.dphi + ((m * G::HO + 2 * yp + oy) * G::WO + 2 * xp + ox) * COUT + co0;
#pragma unroll
          for (int c = 0; c < CO_T; ++c) s[c] = fmaf(acc[g][oy][ox][c] + sdb[co0 + c], __ldg(dp + c), s[c]);
        }
      float* op = a.out + ((b * a.M + m) * NW + wi) * COUT + co0;
#pragma unroll
      for (int c = 0; c < CO_T; ++c) op[c] = 0.25f * s[c];
    }
  }
  }   // rounds
}

// ---- VJP: unpool + mask, kernel / bias gradient partials, delta back-propagation ------------------------------------------------------
struct ConvVjpArgs {
  const float* tin;     // [B, M, HP, WP, COUT]  gradient w.r.t. the pooled stage output
  const float* dphi;    // [M, HO, WO, COUT]
  const float* X;       // [M, HI, WI, CIN]
  const float* W;       // [KK, COUT]            (DGRAD)
  float* gin;           // [B, M, HI, WI, CIN]   gradient w.r.t. the stage input (DGRAD)
  float* part;          // [B, G, KK*COUT + COUT] per-CTA partial kernel / bias gradients
  int M, G, per_cta;    // images per CTA (a multiple of IF)
};

// row stride of the delta image in shared memory: = 4 (mod 32) floats, so that threads on different rows read different banks
__host__ __device__ constexpr int delta_row_stride(int wd_cp) { return wd_cp + (36 - wd_cp % 32) % 32; }

template <class G, int IF, bool DGRAD>
constexpr int vjp_smem_floats(int NT) {
  constexpr int BD = DGRAD ? G::KS - 1 - G::PAD : 0;
  return IF * (G::HO + 2 * BD) * delta_row_stride((G::WO + 2 * BD) * G::CP) + IF * G::CIN * G::PLANE + (DGRAD ? G::KK * G::COUT : 0) +
         NT * 6;
}

// CO_T / CI_T: output / input channels per kernel-gradient thread; RS: row slices of the output image; IF: images in flight
template <class G, int CO_T, int CI_T, int RS, int IF, bool DGRAD, int NT>
__global__ void __launch_bounds__(NT, DGRAD ? 2 : 3) conv5_vjp_kernel(ConvVjpArgs a) {
  constexpr int CIN = G::CIN, COUT = G::COUT, KS = G::KS, CP = G::CP, KK = G::KK, PLANE = G::PLANE, WS = G::WS, HS = G::HS;
  constexpr int HO = G::HO, WO = G::WO, HI = G::HI, WI = G::WI, PAD = G::PAD;
  constexpr int BD = DGRAD ? KS - 1 - PAD : 0, HD = HO + 2 * BD, WD = WO + 2 * BD;
  constexpr int DRS = delta_row_stride(WD * CP);
  constexpr int NCH = COUT / CO_T, NCI = CIN / CI_T, NTW = KS * NCI * NCH * RS, ROWS = HO / RS;
  constexpr int SD = IF * HD * DRS, SX = IF * CIN * PLANE, SW = DGRAD ? KK * COUT : 0;
  static_assert(CIN % CI_T == 0, "input channel group");
  // phase A items: CW channels of one pixel (6 -> three 8-byte loads, else one 16-byte load), stored as CWP floats of the padded row
  constexpr int CW = (COUT % 4 == 0) ? 4 : COUT, CWP = (CW + 3) & ~3, NCHK = COUT / CW;
  static_assert(COUT % CO_T == 0 && HO % RS == 0 && IF * NTW <= NT && NT % NCHK == 0 && NCHK * CWP == CP, "thread roles");
  static_assert(IF * RS * KK * CP <= SD, "the reduction scratch aliases the delta images");
  extern __shared__ __align__(16) float sm[];
  float* sD = sm;                 // [IF][HD][DRS >= WD*CP]  delta w.r.t. the pre-activation, zero border for the transposed conv
  float* sX = sD + SD;            // [IF][CIN][HS][WS]  zero-padded planar input
  float* sW = sX + SX;            // [KK][COUT]
  float* sB = sW + SW;            // [NT][CW]
  const int tid = threadIdx.x;
  const long long b = blockIdx.y;
  const int g = blockIdx.x;
  const int m_begin = g * a.per_cta;
  const int m_end = m_begin + a.per_cta < a.M ? m_begin + a.per_cta : a.M;

  for (int i = tid; i < SD + SX; i += NT) sD[i] = 0.f;
  if (DGRAD)
    for (int i = tid; i < KK * COUT; i += NT) sW[i] = __ldg(a.W + i);

  // kernel-gradient role: (tap row dy, channel group ch, row slice rs, input channel group ci) of image slot img; dy varies
  // fastest: the lanes of a warp then share delta pixels (broadcast) and read neighbouring input rows
  const bool wrole = tid < IF * NTW;
  const int w_dy = tid % KS, w_ch = (tid / KS) % NCH, w_rs = (tid / (KS * NCH)) % RS, w_ci = ((tid / (KS * NCH * RS)) % NCI) * CI_T;
  const int w_img = tid / NTW;
  float acc[CI_T][KS][CO_T];
#pragma unroll
  for (int g = 0; g < CI_T; ++g)
#pragma unroll
    for (int i = 0; i < KS; ++i)
#pragma unroll
      for (int c = 0; c < CO_T; ++c) acc[g][i][c] = 0.f;
  float bsum[CW];
#pragma unroll
  for (int j = 0; j < CW; ++j) bsum[j] = 0.f;
  const int a_chunk = tid % NCHK;
  __syncthreads();

  for (int mb = m_begin; mb < m_end; mb += IF) {
    // phase A: d = phi' * unpool(t_in) / 4 and the planar input images.  One item = CW channels of one pixel; the global loads of
    // a batch of items are issued before any of them is used (the phase is latency-bound otherwise)
    {
      constexpr int NITEM = IF * HO * WO * NCHK;
      constexpr int UA = 4;
      for (int e0 = tid; e0 < NITEM; e0 += NT * UA) {
        float dp[UA][CW], tv[UA][CW];
#pragma unroll
        for (int u = 0; u < UA; ++u) {
          const int e = e0 + u * NT;
          const int q = e / NCHK;                       // chunk = e % NCHK = tid % NCHK (NT % NCHK == 0)
          const int pos = q % (HO * WO), img = q / (HO * WO);
          const int y = pos / WO, x = pos - y * WO;
          const long long m = mb + img;
          const bool ok = e < NITEM && m < m_end;
#pragma unroll
          for (int j = 0; j < CW; ++j) dp[u][j] = tv[u][j] = 0.f;
          if (ok) {
            ldg_vec<CW>(a.dphi + (m * (HO * WO) + pos) * COUT + a_chunk * CW, dp[u]);
            ldg_vec<CW>(a.tin + (((b * a.M + m) * G::HP + (y >> 1)) * G::WP + (x >> 1)) * COUT + a_chunk * CW, tv[u]);
          }
        }
#pragma unroll
        for (int u = 0; u < UA; ++u) {
          const int e = e0 + u * NT;
          if (e < NITEM) {
            const int q = e / NCHK;
            const int pos = q % (HO * WO), img = q / (HO * WO);
            const int y = pos / WO, x = pos - y * WO;
            float v[CWP];
#pragma unroll
            for (int j = 0; j < CWP; ++j) v[j] = 0.f;
#pragma unroll
            for (int j = 0; j < CW; ++j) { v[j] = 0.25f * dp[u][j] * tv[u][j]; bsum[j] += v[j]; }
            float* dst = sD + (img * HD + y + BD) * DRS + (x + BD) * CP + a_chunk * CWP;
#pragma unroll
            for (int j = 0; j < CWP; j += 4) *reinterpret_cast<float4*>(dst + j) = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
          }
        }
      }
      constexpr int PERX = HI * WI * CIN;
      static_assert(PERX % 4 == 0, "images are staged with 16-byte loads");
      constexpr int NX4 = IF * PERX / 4;
      for (int i0 = tid; i0 < NX4; i0 += NT * 2) {
        float4 vx[2];
#pragma unroll
        for (int u = 0; u < 2; ++u) {
          const int i = i0 + u * NT;
          const int img = i / (PERX / 4);
          vx[u] = make_float4(0.f, 0.f, 0.f, 0.f);
          if (i < NX4 && mb + img < m_end) vx[u] = __ldg(reinterpret_cast<const float4*>(a.X + (long long)mb * PERX) + i);
        }
#pragma unroll
        for (int u = 0; u < 2; ++u) {
          const int i = i0 + u * NT;
          if (i < NX4) {
            const float ex[4] = {vx[u].x, vx[u].y, vx[u].z, vx[u].w};
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              const int e = 4 * i + j;
              const int img = e / PERX, r = e - img * PERX;
              const int pix = r / CIN, ci = r - pix * CIN;
              const int y = pix / WI, x = pix - y * WI;
              sX[((img * CIN + ci) * HS + y + PAD) * WS + x + PAD] = ex[j];
            }
          }
        }
      }
    }
    __syncthreads();

    // phase B: gW[(dy, dx, ci), co] += sum_{y, x} X[y + dy, x + dx, ci] * d[y, x, co]
    if (wrole) {
      const float* xr = sX + ((w_img * CIN + w_ci) * HS + w_dy) * WS;
      const float* dr = sD + (w_img * HD + BD) * DRS + BD * CP + w_ch * CO_T;
#pragma unroll 1
      for (int yy = 0; yy < ROWS; ++yy) {
        const int y = w_rs * ROWS + yy;
        float ar[CI_T][WS];
#pragma unroll
        for (int g = 0; g < CI_T; ++g) lds_row(xr + g * PLANE + y * WS, ar[g]);
#pragma unroll
        for (int x = 0; x < WO; ++x) {
          float dv[CO_T];
          lds_vec(dr + y * DRS + x * CP, dv);
#pragma unroll
          for (int g = 0; g < CI_T; ++g)
#pragma unroll
            for (int dx = 0; dx < KS; ++dx)
#pragma unroll
              for (int c = 0; c < CO_T; ++c) acc[g][dx][c] = fmaf(ar[g][x + dx], dv[c], acc[g][dx][c]);
        }
      }
    }

    // phase C: g_in[yi, xi, ci] = sum_{dy, dx, co} d[yi + PAD - dy, xi + PAD - dx, co] * W[(dy, dx, ci), co]
    if constexpr (DGRAD) {
      constexpr int XT = 7, NSEG = WI / XT, TILES = HI * NSEG * (CP / 4);
      static_assert(WI % XT == 0 && CP == COUT, "delta back-propagation tile");
      for (int w = tid; w < IF * TILES; w += NT) {
        const int k4 = w & 3;
        int t = w >> 2;
        const int seg = t % NSEG; t /= NSEG;
        const int yi = t % HI, img = t / HI;
        const int xi0 = seg * XT;
        float o[XT * CIN];
#pragma unroll
        for (int i = 0; i < XT * CIN; ++i) o[i] = 0.f;
#pragma unroll 1
        for (int dy = 0; dy < KS; ++dy) {
          const int yy = yi + PAD - dy;
          if (yy < 0 || yy >= HO) continue;
          // pixel p with tap dx reads padded column xi0 + p - dx + (PAD + BD) = xi0 + p - dx + KS - 1
          const float* dp = sD + (img * HD + yy + BD) * DRS + xi0 * CP + k4 * 4;
          float4 ds[XT + KS - 1];
#pragma unroll
          for (int j = 0; j < XT + KS - 1; ++j) ds[j] = *reinterpret_cast<const float4*>(dp + j * CP);
#pragma unroll
          for (int dx = 0; dx < KS; ++dx)
#pragma unroll
            for (int ci = 0; ci < CIN; ++ci) {
              const float4 wv = *reinterpret_cast<const float4*>(sW + ((dy * KS + dx) * CIN + ci) * COUT + k4 * 4);
#pragma unroll
              for (int p = 0; p < XT; ++p) {
                const float4 dv = ds[p - dx + KS - 1];
                float s = o[p * CIN + ci];
                s = fmaf(dv.x, wv.x, s); s = fmaf(dv.y, wv.y, s); s = fmaf(dv.z, wv.z, s); s = fmaf(dv.w, wv.w, s);
                o[p * CIN + ci] = s;
              }
            }
        }
        // sum the four channel quarters (lanes k4 = 0..3 of one aligned quad), then the quad stores the 42 contiguous floats
#pragma unroll
        for (int i = 0; i < XT * CIN; ++i) {
          o[i] += __shfl_xor_sync(0xffffffffu, o[i], 1);
          o[i] += __shfl_xor_sync(0xffffffffu, o[i], 2);
        }
        const long long m = mb + img;
        if (m < m_end) {
          float* gp = a.gin + (((b * a.M + m) * HI + yi) * WI + xi0) * CIN;
#pragma unroll
          for (int i = 0; i < XT * CIN; ++i)
            if ((i & 3) == k4) gp[i] = o[i];
        }
      }
    }
    __syncthreads();
  }

  // fixed-order reduction over image slots / row slices, then one partial per CTA
  float* red = sD;                 // [IF * RS][KK][CP]
  if (wrole) {
    const int slot = w_img * RS + w_rs;
#pragma unroll
    for (int g = 0; g < CI_T; ++g)
#pragma unroll
      for (int dx = 0; dx < KS; ++dx)
#pragma unroll
        for (int c = 0; c < CO_T; ++c)
          red[(slot * KK + (w_dy * KS + dx) * CIN + w_ci + g) * CP + w_ch * CO_T + c] = acc[g][dx][c];
  }
#pragma unroll
  for (int j = 0; j < CW; ++j) sB[tid * CW + j] = bsum[j];
  __syncthreads();
  float* part = a.part + (b * a.G + g) * (long long)(KK * COUT + COUT);
  for (int o = tid; o < KK * COUT; o += NT) {
    const int k = o / COUT, c = o - k * COUT;
    float s = 0.f;
#pragma unroll 4
    for (int slot = 0; slot < IF * RS; ++slot) s += red[(slot * KK + k) * CP + c];
    part[o] = s;
  }
  if (tid < COUT) {        // channel tid = chunk * CW + j: summed over the threads of that chunk in a fixed order
    const int chunk = tid / CW, j = tid - chunk * CW;
    float s = 0.f;
    for (int t = chunk; t < NT; t += NCHK) s += sB[t * CW + j];
    part[KK * COUT + tid] = s;
  }
}

// out[b][woff + j] = scale * sum_g part[b][g][j] + add_scale * add[b][woff + j]   (j < nW; the bias block follows at boff)
__global__ void __launch_bounds__(256) cnn_part_finish_kernel(const float* __restrict__ part, int G, int nW, int nB, long long B,
                                                              float* __restrict__ out, long long ldo, long long woff, long long boff,
                                                              float scale, const float* __restrict__ add, long long lda,
                                                              float add_scale) {
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= nW + nB) return;
  const long long dst = j < nW ? woff + j : boff + (j - nW);
  for (long long b = blockIdx.y; b < B; b += gridDim.y) {
    const float* p = part + b * G * (long long)(nW + nB) + j;
    float s = 0.f;
    for (int g = 0; g < G; ++g) s += p[(long long)g * (nW + nB)];
    float v = scale * s;
    if (add) v = fmaf(add_scale, __ldg(add + b * lda + dst), v);
    out[b * ldo + dst] = v;
  }
}

// ---- ResNet1M stem: 3x3 conv (3 -> 32 channels, 32x32 image, SAME) JVP + BatchNorm-JVP epilogue ---------------------------------------
// The stem's JVP has no tangent input (its source is data): out = mask * (g * conv(X, dW[b]) + xhat * dscale[b] + dbeta[b]), stored as
// the TF32 (hi, lo) pair the next tcgen05 conv reads.  As a SIMT implicit GEMM (K = 27, N = 32) plus a separate BatchNorm pass it
// moved 2.1 GB of fp32 through HBM twice at 0.2 TB/s (11.9 ms of a 225 ms call at M = 4096); here the image sits in shared memory,
// a thread owns 4 pixels x 8 channels, and the only HBM traffic is xhat / mask in and the pair out.
struct StemJvpArgs {
  const float* X;        // [M, 32, 32, 3]
  const float* V;        // probe block at the kernel offset (probe stride ldv)
  const float* dscale;   // ... at the BatchNorm scale / bias offsets
  const float* dbeta;
  const float* g;        // [32]
  const float* xhat;     // [M, 1024, 32]
  const float* mask;     // [M, 1024, 32] or null
  float* out;            // [B, M, 1024, 32]
  float* out_lo;         // same, or null (plain fp32 tangent)
  long long ldv;
  int M, IMG, ROUNDS;
};

template <int NT>
__global__ void __launch_bounds__(NT, 2) stem3_jvp_bn_kernel(StemJvpArgs a) {
  constexpr int CIN = 3, COUT = 32, KS = 3, H = 32, W = 32, HS = 34, WS = 34, PLANE = HS * WS, KK = KS * KS * CIN, PER = H * W * CIN;
  constexpr int XT = 4, CO_T = 8, NCG = COUT / CO_T, NQ = H * (W / XT);     // pixel quads per image
  extern __shared__ __align__(16) float sm[];
  float* sdW = sm;                         // [KK][COUT]
  float* sP = sdW + KK * COUT;             // g, dscale, dbeta: 3 x [COUT]
  float* sX = sP + 3 * COUT;               // [IMG][CIN][HS][WS]
  const int tid = threadIdx.x;
  const long long b = blockIdx.y;
  for (int i = tid; i < a.IMG * CIN * PLANE; i += NT) sX[i] = 0.f;
  {
    const float* dW = a.V + b * a.ldv;
#pragma unroll 4
    for (int i = tid; i < KK * COUT; i += NT) sdW[i] = __ldg(dW + i);
    if (tid < COUT) {
      sP[tid] = __ldg(a.g + tid);
      sP[COUT + tid] = __ldg(a.dscale + b * a.ldv + tid);
      sP[2 * COUT + tid] = __ldg(a.dbeta + b * a.ldv + tid);
    }
  }
  __syncthreads();
  for (int rd = 0; rd < a.ROUNDS; ++rd) {
    const int m0 = (blockIdx.x * a.ROUNDS + rd) * a.IMG;
    const int nimg = a.M - m0 < a.IMG ? a.M - m0 : a.IMG;
    if (nimg <= 0) break;
    if (rd > 0) __syncthreads();
    {
      const float4* Xg = reinterpret_cast<const float4*>(a.X + (long long)m0 * PER);
      const int total4 = nimg * (PER / 4);
      for (int i0 = tid; i0 < total4; i0 += NT * 4) {
        float4 vx[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) if (i0 + u * NT < total4) vx[u] = __ldg(Xg + i0 + u * NT);
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const int i = i0 + u * NT;
          if (i < total4) {
            const float ex[4] = {vx[u].x, vx[u].y, vx[u].z, vx[u].w};
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              const int e = 4 * i + j;
              const int img = e / PER, r = e - img * PER;
              const int pix = r / CIN, ci = r - pix * CIN;
              const int y = pix / W, x = pix - y * W;
              sX[(img * CIN + ci) * PLANE + (y + 1) * WS + x + 1] = ex[j];
            }
          }
        }
      }
    }
    __syncthreads();
    const int total = nimg * NQ * NCG;
    for (int it = tid; it < total; it += NT) {
      const int cg = it % NCG, w = it / NCG;
      const int img = w / NQ, q = w - img * NQ;
      const int y = q / (W / XT), x0 = (q - y * (W / XT)) * XT;
      const int co0 = cg * CO_T;
      float acc[XT][CO_T];
#pragma unroll
      for (int p = 0; p < XT; ++p)
#pragma unroll
        for (int c = 0; c < CO_T; ++c) acc[p][c] = 0.f;
#pragma unroll
      for (int ci = 0; ci < CIN; ++ci)
#pragma unroll
        for (int dy = 0; dy < KS; ++dy) {
          float xr[XT + KS - 1];
          lds_row(sX + (img * CIN + ci) * PLANE + (y + dy) * WS + x0, xr);
#pragma unroll
          for (int dx = 0; dx < KS; ++dx) {
            float wv[CO_T];
            lds_vec(sdW + ((dy * KS + dx) * CIN + ci) * COUT + co0, wv);
#pragma unroll
            for (int p = 0; p < XT; ++p)
#pragma unroll
              for (int c = 0; c < CO_T; ++c) acc[p][c] = fmaf(xr[p + dx], wv[c], acc[p][c]);
          }
        }
      const long long m = m0 + img;
#pragma unroll
      for (int p = 0; p < XT; ++p) {
        const long long i = ((m * H + y) * W + x0 + p) * COUT + co0;            // index inside one probe's tensor
        const long long o = (b * a.M * (long long)(H * W) * COUT) + i;
        float xh[CO_T], mk[CO_T], hi[CO_T], lo[CO_T];
        const float4 x0v = __ldg(reinterpret_cast<const float4*>(a.xhat + i)), x1v = __ldg(reinterpret_cast<const float4*>(a.xhat + i + 4));
        xh[0] = x0v.x; xh[1] = x0v.y; xh[2] = x0v.z; xh[3] = x0v.w; xh[4] = x1v.x; xh[5] = x1v.y; xh[6] = x1v.z; xh[7] = x1v.w;
        if (a.mask) {
          const float4 m0v = __ldg(reinterpret_cast<const float4*>(a.mask + i)), m1v = __ldg(reinterpret_cast<const float4*>(a.mask + i + 4));
          mk[0] = m0v.x; mk[1] = m0v.y; mk[2] = m0v.z; mk[3] = m0v.w; mk[4] = m1v.x; mk[5] = m1v.y; mk[6] = m1v.z; mk[7] = m1v.w;
        }
#pragma unroll
        for (int c = 0; c < CO_T; ++c) {
          float v = fmaf(sP[co0 + c], acc[p][c], fmaf(xh[c], sP[COUT + co0 + c], sP[2 * COUT + co0 + c]));
          if (a.mask) v *= mk[c];
          if (a.out_lo) {
            const float hh = tf32_round(v);
            hi[c] = hh; lo[c] = tf32_round(v - hh);
          } else {
            hi[c] = v;
          }
        }
        *reinterpret_cast<float4*>(a.out + o) = make_float4(hi[0], hi[1], hi[2], hi[3]);
        *reinterpret_cast<float4*>(a.out + o + 4) = make_float4(hi[4], hi[5], hi[6], hi[7]);
        if (a.out_lo) {
          *reinterpret_cast<float4*>(a.out_lo + o) = make_float4(lo[0], lo[1], lo[2], lo[3]);
          *reinterpret_cast<float4*>(a.out_lo + o + 4) = make_float4(lo[4], lo[5], lo[6], lo[7]);
        }
      }
    }
  }
}

bool fuse_enabled() {
  static const int on = [] {
    const char* e = getenv("LIP_CNN_FUSE");
    return (e && e[0] == '0') ? 0 : 1;
  }();
  return on != 0;
}

template <class G, int CO_T, bool DUAL, int NT>
int launch_jvp(const ConvStage& s, const lip_model* m, const float* V, int64_t ldv, const float* T, float* out, int64_t B, int IMG,
               cudaStream_t st) {
  // rounds per CTA: amortise the kernel staging while keeping >= ~16 CTAs per SM over the probe block (a short last wave)
  int64_t rounds = ceil_div(m->M, IMG) * B / (148 * 16);
  rounds = rounds < 1 ? 1 : (rounds > 8 ? 8 : rounds);
  ConvJvpArgs a;
  a.X = s.Xin; a.T = T; a.W = m->theta + s.woff; a.V = V + s.woff; a.Vb = V + s.boff; a.dphi = s.dphi; a.out = out;
  a.ldv = ldv; a.M = (int)m->M; a.IMG = IMG; a.ROUNDS = (int)rounds;
  const size_t smem = sizeof(float) * (size_t)jvp_smem_floats<G, DUAL>(IMG);
  auto kern = conv5_jvp_pool_kernel<G, CO_T, DUAL, NT>;
  static bool attr_done = false;
  if (!attr_done) {
    LIP_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024));
    attr_done = true;
  }
  const unsigned gx = (unsigned)ceil_div(m->M, IMG * rounds);
  for (int64_t b0 = 0; b0 < B; b0 += 65535) {
    const int64_t nb = B - b0 < 65535 ? B - b0 : 65535;
    ConvJvpArgs c = a;
    c.V += b0 * ldv; c.Vb += b0 * ldv;
    if (T) c.T += b0 * m->M * (int64_t)(G::HI * G::WI * G::CIN);
    c.out += b0 * m->M * (int64_t)(G::HP * G::WP * G::COUT);
    kern<<<dim3(gx, (unsigned)nb), NT, smem, st>>>(c);
    LIP_LAUNCH_CHECK();
  }
  return LIP_OK;
}

template <class G, int CO_T, int CI_T, int RS, int IF, bool DGRAD, int NT>
int launch_vjp(const ConvStage& s, const lip_model* m, const float* tin, float* gin, float* out, int64_t B, float scale,
               const float* add, float add_scale, float* scratch, int64_t scratch_elems, cudaStream_t st) {
  const int64_t M = m->M;
  // enough CTAs for ~16 per SM over the whole probe block (a short last wave), each looping over a multiple of IF images
  int64_t Gn = ceil_div(148 * 16, B);
  const int64_t maxG = ceil_div(M, IF);
  Gn = Gn < 1 ? 1 : (Gn > maxG ? maxG : Gn);
  const int64_t per_cta = ceil_div(ceil_div(M, Gn), IF) * IF;
  Gn = ceil_div(M, per_cta);
  const int nW = G::KK * G::COUT, nB = G::COUT;
  const int64_t Bc_max = scratch_elems / (Gn * (int64_t)(nW + nB));
  if (Bc_max < 1) {
    set_error("conv stage VJP: scratch too small (%lld floats)", (long long)scratch_elems);
    return LIP_ERR_WORKSPACE;
  }
  const size_t smem = sizeof(float) * (size_t)vjp_smem_floats<G, IF, DGRAD>(NT);
  auto kern = conv5_vjp_kernel<G, CO_T, CI_T, RS, IF, DGRAD, NT>;
  static bool attr_done = false;
  if (!attr_done) {
    LIP_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024));
    attr_done = true;
  }
  const int64_t step = Bc_max < 65535 ? Bc_max : 65535;
  for (int64_t b0 = 0; b0 < B; b0 += step) {
    const int64_t nb = B - b0 < step ? B - b0 : step;
    ConvVjpArgs a;
    a.tin = tin + b0 * M * (int64_t)(G::HP * G::WP * G::COUT);
    a.dphi = s.dphi; a.X = s.Xin; a.W = m->theta + s.woff;
    a.gin = gin ? gin + b0 * M * (int64_t)(G::HI * G::WI * G::CIN) : nullptr;
    a.part = scratch; a.M = (int)M; a.G = (int)Gn; a.per_cta = (int)per_cta;
    kern<<<dim3((unsigned)Gn, (unsigned)nb), NT, smem, st>>>(a);
    LIP_LAUNCH_CHECK();
    cnn_part_finish_kernel<<<dim3((unsigned)ceil_div(nW + nB, 256), (unsigned)nb), 256, 0, st>>>(
        scratch, (int)Gn, nW, nB, nb, out + b0 * m->D, m->D, s.woff, s.boff, scale, add ? add + b0 * m->D : nullptr, m->D, add_scale);
    LIP_LAUNCH_CHECK();
  }
  return LIP_OK;
}

}  // namespace

bool resnet_stem_fusable(const ConvBN& u) {
  return fuse_enabled() && u.src == -2 && u.skip < 0 && u.cin == 3 && u.cout == 32 && u.kh == 3 && u.kw == 3 && u.stride == 1 &&
         u.pad_h == 1 && u.pad_w == 1 && u.Hi == 32 && u.Wi == 32 && u.Ho == 32 && u.Wo == 32 && u.Xin && u.xhat && u.g;
}

int resnet_stem_jvp(const lip_model* m, const ConvBN& u, const float* V, int64_t ldv, float* out, float* out_lo, int64_t B,
                    cudaStream_t st) {
  constexpr int NT = 256, IMG = 2;
  StemJvpArgs a;
  a.X = u.Xin; a.V = V + u.woff; a.dscale = V + u.scale_off; a.dbeta = V + u.beta_off; a.g = u.g; a.xhat = u.xhat; a.mask = u.mask;
  a.ldv = ldv; a.M = (int)m->M; a.IMG = IMG;
  int64_t rounds = ceil_div(m->M, IMG) * B / (148 * 16);
  rounds = rounds < 1 ? 1 : (rounds > 16 ? 16 : rounds);
  a.ROUNDS = (int)rounds;
  const size_t smem = sizeof(float) * (size_t)(27 * 32 + 3 * 32 + IMG * 3 * 34 * 34);
  const unsigned gx = (unsigned)ceil_div(m->M, IMG * rounds);
  const int64_t per_probe = m->M * (int64_t)(32 * 32 * 32);
  for (int64_t b0 = 0; b0 < B; b0 += 65535) {
    const int64_t nb = B - b0 < 65535 ? B - b0 : 65535;
    StemJvpArgs c = a;
    c.V += b0 * ldv; c.dscale += b0 * ldv; c.dbeta += b0 * ldv;
    c.out = out + b0 * per_probe;
    c.out_lo = out_lo ? out_lo + b0 * per_probe : nullptr;
    stem3_jvp_bn_kernel<NT><<<dim3(gx, (unsigned)nb), NT, smem, st>>>(c);
    LIP_LAUNCH_CHECK();
  }
  return LIP_OK;
}

bool cnn_stage_fusable(const lip_model* m, int i) {
  if (!fuse_enabled() || i < 0 || i >= (int)m->CS.size()) return false;
  const ConvStage& s = m->CS[i];
  if (s.Xin == nullptr || s.dphi == nullptr) return false;
  if (i == 0) return GeoC1::match(s);
  return GeoC2::match(s);
}

int cnn_fused_jvp(const lip_model* m, int i, const float* V, int64_t ldv, const float* T, float* out, int64_t B, cudaStream_t st) {
  const ConvStage& s = m->CS[i];
  if (i == 0) return launch_jvp<GeoC1, 6, false, 224>(s, m, V, ldv, nullptr, out, B, 16, st);     // 8 pairs x 196 windows = 7 x 224
  return launch_jvp<GeoC2, 8, true, 256>(s, m, V, ldv, T, out, B, 10, st);     // 5 pairs x 25 windows x 2 groups = 250 items of 256
}

int cnn_fused_vjp(const lip_model* m, int i, const float* tin, float* gin, float* out, int64_t B, float scale, const float* add,
                  float add_scale, float* scratch, int64_t scratch_elems, cudaStream_t st) {
  const ConvStage& s = m->CS[i];
  if (i == 0)
    return launch_vjp<GeoC1, 6, 1, 28, 2, false, 288>(s, m, tin, nullptr, out, B, scale, add, add_scale, scratch, scratch_elems, st);
  // one input channel per kernel-gradient thread and no row slices measured faster than 2 x 2 (register pressure): 6.22 vs 6.46 ms / step
  return launch_vjp<GeoC2, 4, 1, 1, 2, true, 256>(s, m, tin, gin, out, B, scale, add, add_scale, scratch, scratch_elems, st);
}

}  // namespace lip
