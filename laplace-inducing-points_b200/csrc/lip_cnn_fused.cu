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
  static constexpr int HS = HI + 2 * PAD, WS = (WI + 2 * PAD + 1) & ~1;     // zero-padded planar input image in shared memory
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
  int M, IMG;           // images per CTA
};

template <class G, bool DUAL>
constexpr int jvp_smem_floats(int IMG) {
  return G::KK * G::CP * (DUAL ? 2 : 1) + 16 + IMG * G::CIN * G::PLANE * (DUAL ? 2 : 1);
}

template <class G, int CO_T, bool DUAL, int NT>
__global__ void __launch_bounds__(NT) conv5_jvp_pool_kernel(ConvJvpArgs a) {
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
  const int m0 = blockIdx.x * a.IMG;
  const int nimg = a.M - m0 < a.IMG ? a.M - m0 : a.IMG;

  for (int i = tid; i < a.IMG * CIN * PLANE * (DUAL ? 2 : 1); i += NT) sX[i] = 0.f;
  {
    const float* dW = a.V + b * a.ldv;
    for (int i = tid; i < KK * CP; i += NT) {
      const int k = i / CP, c = i - k * CP;
      sdW[i] = c < COUT ? __ldg(dW + k * COUT + c) : 0.f;
      if (DUAL) sW[i] = c < COUT ? __ldg(a.W + k * COUT + c) : 0.f;
    }
    if (tid < 16) sdb[tid] = tid < COUT ? __ldg(a.Vb + b * a.ldv + tid) : 0.f;
  }
  __syncthreads();
  {
    constexpr int PER = G::HI * G::WI * CIN;
    const float* Xg = a.X + (long long)m0 * PER;
    const float* Tg = DUAL ? a.T + (b * a.M + m0) * PER : nullptr;
    for (int i = tid; i < nimg * PER; i += NT) {
      const int img = i / PER, r = i - img * PER;
      const int pix = r / CIN, ci = r - pix * CIN;
      const int y = pix / G::WI, x = pix - y * G::WI;
      const int o = (img * CIN + ci) * PLANE + (y + G::PAD) * WS + x + G::PAD;
      sX[o] = __ldg(Xg + i);
      if (DUAL) sT[o] = __ldg(Tg + i);
    }
  }
  __syncthreads();

  const int total = nimg * NW * NCG;
  for (int it = tid; it < total; it += NT) {
    const int cg = it % NCG, w = it / NCG;
    const int img = w / NW, wi = w - img * NW;
    const int yp = wi / G::WP, xp = wi - yp * G::WP;
    const int co0 = cg * CO_T;
    float acc[2][2][CO_T];
#pragma unroll
    for (int i = 0; i < 2; ++i)
#pragma unroll
      for (int j = 0; j < 2; ++j)
#pragma unroll
        for (int c = 0; c < CO_T; ++c) acc[i][j][c] = 0.f;
    const int woff = img * CIN * PLANE + 2 * yp * WS + 2 * xp;
#pragma unroll 1
    for (int ci = 0; ci < CIN; ++ci) {
      const float* px = sX + woff + ci * PLANE;
      const float* pt = sT + woff + ci * PLANE;
      float ra[KS + 1], rb[KS + 1], ta[KS + 1], tb[KS + 1];
      lds_row(px, ra);
      if (DUAL) lds_row(pt, ta);
#pragma unroll
      for (int dy = 0; dy < KS; ++dy) {
        lds_row(px + (dy + 1) * WS, rb);
        if (DUAL) lds_row(pt + (dy + 1) * WS, tb);
#pragma unroll
        for (int dx = 0; dx < KS; ++dx) {
          const int k = ((dy * KS + dx) * CIN + ci) * CP + co0;
          float wv[CO_T];
          lds_vec(sdW + k, wv);
#pragma unroll
          for (int c = 0; c < CO_T; ++c) {
            acc[0][0][c] = fmaf(ra[dx], wv[c], acc[0][0][c]);
            acc[0][1][c] = fmaf(ra[dx + 1], wv[c], acc[0][1][c]);
            acc[1][0][c] = fmaf(rb[dx], wv[c], acc[1][0][c]);
            acc[1][1][c] = fmaf(rb[dx + 1], wv[c], acc[1][1][c]);
          }
          if (DUAL) {
            lds_vec(sW + k, wv);
#pragma unroll
            for (int c = 0; c < CO_T; ++c) {
              acc[0][0][c] = fmaf(ta[dx], wv[c], acc[0][0][c]);
              acc[0][1][c] = fmaf(ta[dx + 1], wv[c], acc[0][1][c]);
              acc[1][0][c] = fmaf(tb[dx], wv[c], acc[1][0][c]);
              acc[1][1][c] = fmaf(tb[dx + 1], wv[c], acc[1][1][c]);
            }
          }
        }
#pragma unroll
        for (int j = 0; j <= KS; ++j) { ra[j] = rb[j]; if (DUAL) ta[j] = tb[j]; }
      }
    }
    // bias tangent, activation mask, 2x2 mean
    const long long m = m0 + img;
    float s[CO_T];
#pragma unroll
    for (int c = 0; c < CO_T; ++c) s[c] = 0.f;
#pragma unroll
    for (int oy = 0; oy < 2; ++oy)
#pragma unroll
      for (int ox = 0; ox < 2; ++ox) {
        const float* dp = a.dphi + ((m * G::HO + 2 * yp + oy) * G::WO + 2 * xp + ox) * COUT + co0;
#pragma unroll
        for (int c = 0; c < CO_T; ++c) s[c] = fmaf(acc[oy][ox][c] + sdb[co0 + c], __ldg(dp + c), s[c]);
      }
    float* op = a.out + ((b * a.M + m) * NW + wi) * COUT + co0;
#pragma unroll
    for (int c = 0; c < CO_T; ++c) op[c] = 0.25f * s[c];
  }
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

template <class G, int IF, bool DGRAD>
constexpr int vjp_smem_floats(int NT) {
  constexpr int BD = DGRAD ? G::KS - 1 - G::PAD : 0;
  return IF * (G::HO + 2 * BD) * (G::WO + 2 * BD) * G::CP + IF * G::CIN * G::PLANE + (DGRAD ? G::KK * G::COUT : 0) + NT;
}

// CO_T: channels per kernel-gradient thread; RS: row slices of the output image per kernel-gradient thread group; IF: images in flight
template <class G, int CO_T, int RS, int IF, bool DGRAD, int NT>
__global__ void __launch_bounds__(NT, 2) conv5_vjp_kernel(ConvVjpArgs a) {
  constexpr int CIN = G::CIN, COUT = G::COUT, KS = G::KS, CP = G::CP, KK = G::KK, PLANE = G::PLANE, WS = G::WS, HS = G::HS;
  constexpr int HO = G::HO, WO = G::WO, HI = G::HI, WI = G::WI, PAD = G::PAD;
  constexpr int BD = DGRAD ? KS - 1 - PAD : 0, HD = HO + 2 * BD, WD = WO + 2 * BD;
  constexpr int NCH = COUT / CO_T, NTW = KS * CIN * NCH * RS, ROWS = HO / RS;
  constexpr int SD = IF * HD * WD * CP, SX = IF * CIN * PLANE, SW = DGRAD ? KK * COUT : 0;
  static_assert(COUT % CO_T == 0 && HO % RS == 0 && IF * NTW <= NT && NT % CP == 0, "thread roles");
  static_assert(IF * RS * KK * CP <= SD, "the reduction scratch aliases the delta images");
  extern __shared__ __align__(16) float sm[];
  float* sD = sm;                 // [IF][HD][WD][CP]   delta w.r.t. the pre-activation, zero border for the transposed conv
  float* sX = sD + SD;            // [IF][CIN][HS][WS]  zero-padded planar input
  float* sW = sX + SX;            // [KK][COUT]
  float* sB = sW + SW;            // [NT]
  const int tid = threadIdx.x;
  const long long b = blockIdx.y;
  const int g = blockIdx.x;
  const int m_begin = g * a.per_cta;
  const int m_end = m_begin + a.per_cta < a.M ? m_begin + a.per_cta : a.M;

  for (int i = tid; i < SD + SX; i += NT) sD[i] = 0.f;
  if (DGRAD)
    for (int i = tid; i < KK * COUT; i += NT) sW[i] = __ldg(a.W + i);

  // kernel-gradient role: (tap row dy, input channel ci, channel group ch, row slice rs) of image slot img
  const bool wrole = tid < IF * NTW;
  const int w_ch = tid % NCH, w_rs = (tid / NCH) % RS, w_ci = (tid / (NCH * RS)) % CIN, w_dy = (tid / (NCH * RS * CIN)) % KS;
  const int w_img = tid / NTW;
  float acc[KS][CO_T];
#pragma unroll
  for (int i = 0; i < KS; ++i)
#pragma unroll
    for (int c = 0; c < CO_T; ++c) acc[i][c] = 0.f;
  float bsum = 0.f;
  __syncthreads();

  for (int mb = m_begin; mb < m_end; mb += IF) {
    // phase A: d = phi' * unpool(t_in) / 4 and the planar input images
    for (int e = tid; e < IF * HO * WO * CP; e += NT) {
      const int c = e % CP, q = e / CP;
      const int pos = q % (HO * WO), img = q / (HO * WO);
      const int y = pos / WO, x = pos - y * WO;
      const long long m = mb + img;
      float v = 0.f;
      if (c < COUT && m < m_end)
        v = 0.25f * __ldg(a.dphi + (m * (HO * WO) + pos) * COUT + c) *
            __ldg(a.tin + (((b * a.M + m) * G::HP + (y >> 1)) * G::WP + (x >> 1)) * COUT + c);
      sD[((img * HD + y + BD) * WD + x + BD) * CP + c] = v;
      bsum += v;
    }
    for (int e = tid; e < IF * HI * WI * CIN; e += NT) {
      const int ci = e % CIN, q = e / CIN;
      const int pix = q % (HI * WI), img = q / (HI * WI);
      const int y = pix / WI, x = pix - y * WI;
      const long long m = mb + img;
      sX[((img * CIN + ci) * HS + y + PAD) * WS + x + PAD] = m < m_end ? __ldg(a.X + (m * (HI * WI) + pix) * CIN + ci) : 0.f;
    }
    __syncthreads();

    // phase B: gW[(dy, dx, ci), co] += sum_{y, x} X[y + dy, x + dx, ci] * d[y, x, co]
    if (wrole) {
      const float* xr = sX + ((w_img * CIN + w_ci) * HS + w_dy) * WS;
      const float* dr = sD + ((w_img * HD + BD) * WD + BD) * CP + w_ch * CO_T;
#pragma unroll 1
      for (int yy = 0; yy < ROWS; ++yy) {
        const int y = w_rs * ROWS + yy;
        float ar[WS];
        lds_row(xr + y * WS, ar);
#pragma unroll
        for (int x = 0; x < WO; ++x) {
          float dv[CO_T];
          lds_vec(dr + (y * WD + x) * CP, dv);
#pragma unroll
          for (int dx = 0; dx < KS; ++dx)
#pragma unroll
            for (int c = 0; c < CO_T; ++c) acc[dx][c] = fmaf(ar[x + dx], dv[c], acc[dx][c]);
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
          const float* dp = sD + ((img * HD + yy + BD) * WD + xi0) * CP + k4 * 4;
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
    for (int dx = 0; dx < KS; ++dx)
#pragma unroll
      for (int c = 0; c < CO_T; ++c)
        red[(slot * KK + (w_dy * KS + dx) * CIN + w_ci) * CP + w_ch * CO_T + c] = acc[dx][c];
  }
  sB[tid] = bsum;
  __syncthreads();
  float* part = a.part + (b * a.G + g) * (long long)(KK * COUT + COUT);
  for (int o = tid; o < KK * COUT; o += NT) {
    const int k = o / COUT, c = o - k * COUT;
    float s = 0.f;
#pragma unroll 4
    for (int slot = 0; slot < IF * RS; ++slot) s += red[(slot * KK + k) * CP + c];
    part[o] = s;
  }
  if (tid < COUT) {
    float s = 0.f;
    for (int t = tid; t < NT; t += CP) s += sB[t];
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
  ConvJvpArgs a;
  a.X = s.Xin; a.T = T; a.W = m->theta + s.woff; a.V = V + s.woff; a.Vb = V + s.boff; a.dphi = s.dphi; a.out = out;
  a.ldv = ldv; a.M = (int)m->M; a.IMG = IMG;
  const size_t smem = sizeof(float) * (size_t)jvp_smem_floats<G, DUAL>(IMG);
  auto kern = conv5_jvp_pool_kernel<G, CO_T, DUAL, NT>;
  static bool attr_done = false;
  if (!attr_done) {
    LIP_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024));
    attr_done = true;
  }
  const unsigned gx = (unsigned)ceil_div(m->M, IMG);
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

template <class G, int CO_T, int RS, int IF, bool DGRAD, int NT>
int launch_vjp(const ConvStage& s, const lip_model* m, const float* tin, float* gin, float* out, int64_t B, float scale,
               const float* add, float add_scale, float* scratch, int64_t scratch_elems, cudaStream_t st) {
  const int64_t M = m->M;
  // enough CTAs for ~8 per SM over the whole probe block, each looping over a multiple of IF images
  int64_t Gn = ceil_div(148 * 8, B);
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
  auto kern = conv5_vjp_kernel<G, CO_T, RS, IF, DGRAD, NT>;
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

bool cnn_stage_fusable(const lip_model* m, int i) {
  if (!fuse_enabled() || i < 0 || i >= (int)m->CS.size()) return false;
  const ConvStage& s = m->CS[i];
  if (s.Xin == nullptr || s.dphi == nullptr) return false;
  if (i == 0) return GeoC1::match(s);
  return GeoC2::match(s);
}

int cnn_fused_jvp(const lip_model* m, int i, const float* V, int64_t ldv, const float* T, float* out, int64_t B, cudaStream_t st) {
  const ConvStage& s = m->CS[i];
  if (i == 0) return launch_jvp<GeoC1, 6, false, 224>(s, m, V, ldv, nullptr, out, B, 8, st);
  return launch_jvp<GeoC2, 8, true, 224>(s, m, V, ldv, T, out, B, 4, st);
}

int cnn_fused_vjp(const lip_model* m, int i, const float* tin, float* gin, float* out, int64_t B, float scale, const float* add,
                  float add_scale, float* scratch, int64_t scratch_elems, cudaStream_t st) {
  const ConvStage& s = m->CS[i];
  if (i == 0) return launch_vjp<GeoC1, 6, 28, 2, false, 288>(s, m, tin, nullptr, out, B, scale, add, add_scale, scratch, scratch_elems, st);
  return launch_vjp<GeoC2, 4, 1, 2, true, 256>(s, m, tin, gin, out, B, scale, add, add_scale, scratch, scratch_elems, st);
}

}  // namespace lip
