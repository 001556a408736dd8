// lip_gemm_simt.cu — exact-fp32 SIMT strided batched GEMM with the hot path's fused epilogue.
//
// This is the any-shape path: toy models (D <= 2274, launch-bound), the K=10 logit layer, the forward
// pass at bind time, and the reference result the tcgen05 3xTF32 kernel is self-tested against.
// 128x128x16 CTA tile, 256 threads, 8x8 register micro-tile (split 4+4 so shared-memory reads are
// conflict-free float4), register-prefetch double buffering.
#include <stdlib.h>

#include "lip_common.cuh"

namespace lip {

namespace {

constexpr int BM = 128, BN = 128, BK = 16, NT = 256, PAD = 4;

struct DevOperand {
  const float* ptr;
  long long sz, s0, s1;
  ConvGather conv;
  // B operands only, nin > 0: the column index is a folded (batch entry, column) pair, n' = q * nin + r  ->  q * fold_sz + r * s1
  int nin;
  FastDiv dnin;
  long long fold_sz;
};

// offset of the image element behind A[m][k] of a gathered operand; false = zero padding
__device__ __forceinline__ bool conv_offset(const ConvGather& c, uint32_t m, uint32_t k, long long* off) {
  if (c.mode == 3) {
    uint32_t t, co, dx, dy, xi, yi, mz;
    c.dC.divmod(k, t, co);
    c.dkw.divmod(t, dy, dx);
    c.dWi.divmod(m, t, xi);
    c.dHi.divmod(t, mz, yi);
    const int ys = (int)yi + c.pad_h - (int)dy, xs = (int)xi + c.pad_w - (int)dx;
    if (ys < 0 || xs < 0) return false;
    int y = ys, x = xs;
    if (c.stride != 1) {
      if ((ys % c.stride) | (xs % c.stride)) return false;
      y = ys / c.stride; x = xs / c.stride;
    }
    if (y >= c.Ho || x >= c.Wo) return false;
    *off = (((long long)mz * c.Ho + y) * c.Wo + x) * c.C + co;
    return true;
  }
  const uint32_t r = c.mode == 1 ? m : k, kc = c.mode == 1 ? k : m;
  uint32_t t, ch, dx, dy, x, y, mz;
  c.dC.divmod(kc, t, ch);
  c.dkw.divmod(t, dy, dx);
  c.dWo.divmod(r, t, x);
  c.dHo.divmod(t, mz, y);
  const int yi = (int)y * c.stride + (int)dy - c.pad_h, xi = (int)x * c.stride + (int)dx - c.pad_w;
  if (yi < 0 || yi >= c.Hi || xi < 0 || xi >= c.Wi) return false;
  *off = (((long long)mz * c.Hi + yi) * c.Wi + xi) * c.C + ch;
  return true;
}

struct DevGemm {
  int M, N, K1, K2;
  DevOperand A1, B1, A2, B2;
  float* C;
  long long c_sz, c_sm;
  float scale;
  const float* bias; long long bias_sz;
  const float* mask; long long mask_sm;
  const float* add;  long long add_sz; float add_scale;
  int act;
  float* dphi_out;
  float* C_lo;
  int nin;             // > 0: batch folded into N (see gemm_simt): C / add column n' = q * nin + r lives at q * c_sz (add_sz) + m * c_sm + r
  FastDiv dnin;
  int ksplit;          // > 1: blockIdx.y = m_tile * ksplit + slice; raw partial tiles go to `part`
  float* part;         // [slice][z][M][N]
  long long part_sz;   // stride between slices = batch * M * N
};

// offset of output element (z, m, n) in an array with batch stride zs (C: c_sz, add: add_sz)
__device__ __forceinline__ long long out_offset(const DevGemm& g, long long z, int m, int n, long long zs) {
  if (g.nin) {
    uint32_t q, r;
    g.dnin.divmod((uint32_t)n, q, r);
    return (long long)q * zs + (long long)m * g.c_sm + r;
  }
  return z * zs + (long long)m * g.c_sm + n;
}

// KC: the contraction index is the contiguous one for this operand; ROWS: tile extent of the other; GATHER: the A operand may
// be an implicit-GEMM patch gather (compiled out of the plain kernels: the divmod chain costs registers / occupancy)
template <bool KC, int ROWS, bool GATHER>
__device__ __forceinline__ void load_tile(const DevOperand& op, long long zoff, int row0, int rows, int k0, int K,
                                          bool is_a, float (&reg)[ROWS * BK / NT]) {
  // A tile: [ROWS rows(m)] x [BK k]; B tile: [BK k] x [ROWS cols(n)].  "row" below = the non-k index.
  const int tid = threadIdx.x;
#pragma unroll
  for (int i = 0; i < ROWS * BK / NT; ++i) {
    int idx = tid + NT * i;
    int k, r;
    if (KC) { k = idx % BK; r = idx / BK; } else { r = idx % ROWS; k = idx / ROWS; }
    int gr = row0 + r, gk = k0 + k;
    float v = 0.f;
    if (gr < rows && gk < K) {
      if (GATHER && is_a && op.conv.mode) {
        long long off;
        if (conv_offset(op.conv, (uint32_t)gr, (uint32_t)gk, &off)) v = __ldg(op.ptr + zoff + off);
      } else {
        long long off;
        if (is_a) {
          off = (long long)gr * op.s0 + (long long)gk * op.s1;
        } else if (op.nin) {
          uint32_t q, r;
          op.dnin.divmod((uint32_t)gr, q, r);
          off = (long long)gk * op.s0 + (long long)q * op.fold_sz + (long long)r * op.s1;
        } else {
          off = (long long)gk * op.s0 + (long long)gr * op.s1;
        }
        v = __ldg(op.ptr + zoff + off);
      }
    }
    reg[i] = v;
  }
}

// Per-thread state of a gathered A tile (128 rows x BK, 8 elements per thread).  The decomposition of the part of the index that
// does not change over the K loop is hoisted: in the k-contiguous layout (modes 1, 3) a thread's 8 elements sit in 8 fixed tile
// rows and share one k, so the rows' image coordinates are computed once per tile and each k-tile costs a single
// (dy, dx, channel) decomposition; in the m-contiguous layout (mode 2) the thread's patch column kc is fixed instead.
template <bool KC>
struct GatherTile {
  int y0[8], x0[8];           // KC: per row  (mode 1: y*stride - pad, x*stride - pad;  mode 3: yi + pad, xi + pad)
  int base[8];                // KC: per row image offset of (mz, 0, 0, 0) (< 2^31: one probe's image);  !KC: unused
  int dy, dx, ch;             // !KC: the fixed patch column
  bool rowok[8], kcok;

  __device__ __forceinline__ void init(const ConvGather& c, int m0, int M) {
    const int tid = threadIdx.x;
    if (KC) {
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const uint32_t gr = (uint32_t)(m0 + (tid + NT * i) / BK);
        rowok[i] = (int)gr < M;
        uint32_t t, x, y, mz;
        if (c.mode == 3) {
          c.dWi.divmod(gr, t, x);
          c.dHi.divmod(t, mz, y);
          y0[i] = (int)y + c.pad_h; x0[i] = (int)x + c.pad_w;
          base[i] = (int)(mz * (uint32_t)(c.Ho * c.Wo * c.C));
        } else {
          c.dWo.divmod(gr, t, x);
          c.dHo.divmod(t, mz, y);
          y0[i] = (int)y * c.stride - c.pad_h; x0[i] = (int)x * c.stride - c.pad_w;
          base[i] = (int)(mz * (uint32_t)(c.Hi * c.Wi * c.C));
        }
      }
    } else {
      const uint32_t kc = (uint32_t)(m0 + tid % BM);
      kcok = (int)kc < M;
      uint32_t t, c_, dx_, dy_;
      c.dC.divmod(kc, t, c_);
      c.dkw.divmod(t, dy_, dx_);
      dy = (int)dy_; dx = (int)dx_; ch = (int)c_;
    }
  }

  __device__ __forceinline__ void load(const DevOperand& op, long long zoff, int k0, int K, float (&reg)[8]) const {
    const ConvGather& c = op.conv;
    const int tid = threadIdx.x;
    const float* p = op.ptr + zoff;
    if (KC) {
      const uint32_t gk = (uint32_t)(k0 + tid % BK);
      uint32_t t, c_, dx_, dy_;
      c.dC.divmod(gk, t, c_);
      c.dkw.divmod(t, dy_, dx_);
      const bool kok = (int)gk < K;
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        float v = 0.f;
        if (kok && rowok[i]) {
          if (c.mode == 3) {
            int ys = y0[i] - (int)dy_, xs = x0[i] - (int)dx_;
            bool ok = ys >= 0 && xs >= 0;
            if (ok && c.stride != 1) { ok = (ys % c.stride == 0) && (xs % c.stride == 0); ys /= c.stride; xs /= c.stride; }
            if (ok && ys < c.Ho && xs < c.Wo) v = __ldg(p + (base[i] + (ys * c.Wo + xs) * c.C + (int)c_));
          } else {
            const int yi = y0[i] + (int)dy_, xi = x0[i] + (int)dx_;
            if (yi >= 0 && yi < c.Hi && xi >= 0 && xi < c.Wi) v = __ldg(p + (base[i] + (yi * c.Wi + xi) * c.C + (int)c_));
          }
        }
        reg[i] = v;
      }
    } else {
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const uint32_t r = (uint32_t)(k0 + (tid + NT * i) / BM);
        float v = 0.f;
        if (kcok && (int)r < K) {
          uint32_t t, x, y, mz;
          c.dWo.divmod(r, t, x);
          c.dHo.divmod(t, mz, y);
          const int yi = (int)y * c.stride + dy - c.pad_h, xi = (int)x * c.stride + dx - c.pad_w;
          if (yi >= 0 && yi < c.Hi && xi >= 0 && xi < c.Wi) v = __ldg(p + (((long long)mz * c.Hi + yi) * c.Wi + xi) * c.C + ch);
        }
        reg[i] = v;
      }
    }
  }
};

template <bool KC, int ROWS>
__device__ __forceinline__ void store_tile(float (*S)[ROWS + PAD], const float (&reg)[ROWS * BK / NT]) {
  const int tid = threadIdx.x;
#pragma unroll
  for (int i = 0; i < ROWS * BK / NT; ++i) {
    int idx = tid + NT * i;
    int k, r;
    if (KC) { k = idx % BK; r = idx / BK; } else { r = idx % ROWS; k = idx / ROWS; }
    S[k][r] = reg[i];
  }
}

// BNT: tile width (128, or 64 / 32 for narrow outputs such as 32- and 64-channel convs); NJ = BNT / 16 columns per thread
template <bool A_KC, bool B_KC, int BNT, bool GATHER>
__global__ void __launch_bounds__(NT, 2) gemm_simt_kernel(DevGemm g) {
  constexpr int NJ = BNT / 16;
  __shared__ __align__(16) float As[2][BK][BM + PAD];
  __shared__ __align__(16) float Bs[2][BK][BNT + PAD];

  const int z = blockIdx.z;
  const int ks = g.ksplit > 1 ? g.ksplit : 1;
  const int slice = blockIdx.y % ks;
  const int m0 = (blockIdx.y / ks) * BM, n0 = blockIdx.x * BNT;
  const int tx = threadIdx.x % 16, ty = threadIdx.x / 16;
  // column j of this thread's micro-tile -> column of the tile
  auto col_of = [&](int j) { return NJ == 8 ? (j < 4 ? tx * 4 + j : 64 + tx * 4 + (j - 4)) : tx * NJ + j; };

  float acc[8][NJ];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < NJ; ++j) acc[i][j] = 0.f;

  const int npairs = g.A2.ptr ? 2 : 1;
  for (int pair = 0; pair < npairs; ++pair) {
    const DevOperand& A = pair ? g.A2 : g.A1;
    const DevOperand& B = pair ? g.B2 : g.B1;
    const int K = pair ? g.K2 : g.K1;
    const long long za = (long long)z * A.sz, zb = (long long)z * B.sz;
    const int nk_all = (K + BK - 1) / BK;
    const int kt0 = (int)((long long)nk_all * slice / ks), nk = (int)((long long)nk_all * (slice + 1) / ks);
    if (kt0 >= nk) continue;
    float ra[8], rb[BNT * BK / NT];
    GatherTile<A_KC> gt;
    const bool gathered = GATHER && A.conv.mode != 0;
    if (gathered) gt.init(A.conv, m0, g.M);
    if (gathered) gt.load(A, za, kt0 * BK, K, ra);
    else load_tile<A_KC, BM, false>(A, za, m0, g.M, kt0 * BK, K, true, ra);
    load_tile<B_KC, BNT, false>(B, zb, n0, g.N, kt0 * BK, K, false, rb);
    store_tile<A_KC, BM>(As[kt0 & 1], ra);
    store_tile<B_KC, BNT>(Bs[kt0 & 1], rb);
    __syncthreads();
    for (int kt = kt0; kt < nk; ++kt) {
      const int cur = kt & 1;
      if (kt + 1 < nk) {
        if (gathered) gt.load(A, za, (kt + 1) * BK, K, ra);
        else load_tile<A_KC, BM, false>(A, za, m0, g.M, (kt + 1) * BK, K, true, ra);
        load_tile<B_KC, BNT, false>(B, zb, n0, g.N, (kt + 1) * BK, K, false, rb);
      }
#pragma unroll
      for (int k = 0; k < BK; ++k) {
        float4 a0 = *reinterpret_cast<const float4*>(&As[cur][k][ty * 4]);
        float4 a1 = *reinterpret_cast<const float4*>(&As[cur][k][64 + ty * 4]);
        float a[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
        float b[NJ];
        if (NJ == 8) {
          float4 b0 = *reinterpret_cast<const float4*>(&Bs[cur][k][tx * 4]);
          float4 b1 = *reinterpret_cast<const float4*>(&Bs[cur][k][64 + tx * 4]);
          b[0] = b0.x; b[1] = b0.y; b[2] = b0.z; b[3] = b0.w;
          b[NJ - 4] = b1.x; b[NJ - 3] = b1.y; b[NJ - 2] = b1.z; b[NJ - 1] = b1.w;
        } else if (NJ == 4) {
          float4 b0 = *reinterpret_cast<const float4*>(&Bs[cur][k][tx * 4]);
          b[0] = b0.x; b[1] = b0.y; b[2] = b0.z; b[3] = b0.w;
        } else {
          float2 b0 = *reinterpret_cast<const float2*>(&Bs[cur][k][tx * 2]);
          b[0] = b0.x; b[1] = b0.y;
        }
#pragma unroll
        for (int i = 0; i < 8; ++i)
#pragma unroll
          for (int j = 0; j < NJ; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
      }
      if (kt + 1 < nk) {
        store_tile<A_KC, BM>(As[cur ^ 1], ra);
        store_tile<B_KC, BNT>(Bs[cur ^ 1], rb);
      }
      __syncthreads();
    }
  }

  if (ks > 1) {   // split-K: raw partial tile, reduced (with the epilogue) by splitk_reduce_kernel
    float* pt = g.part + (long long)slice * g.part_sz + (long long)z * g.M * g.N;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int m = m0 + (i < 4 ? ty * 4 + i : 64 + ty * 4 + (i - 4));
      if (m >= g.M) continue;
#pragma unroll
      for (int j = 0; j < NJ; ++j) {
        const int n = n0 + col_of(j);
        if (n < g.N) pt[(long long)m * g.N + n] = acc[i][j];
      }
    }
    return;
  }
  // ---- fused epilogue ----
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int m = m0 + (i < 4 ? ty * 4 + i : 64 + ty * 4 + (i - 4));
    if (m >= g.M) continue;
#pragma unroll
    for (int j = 0; j < NJ; ++j) {
      const int n = n0 + col_of(j);
      if (n >= g.N) continue;
      float v = g.scale * acc[i][j];
      if (g.bias) v += __ldg(g.bias + (long long)z * g.bias_sz + n);
      const long long co = out_offset(g, z, m, n, g.c_sz);
      if (g.act >= 0) {
        float d;
        v = act_apply(g.act, v, &d);
        if (g.dphi_out) g.dphi_out[co] = d;
      }
      if (g.mask) v *= __ldg(g.mask + (long long)m * g.mask_sm + n);
      if (g.add) v += g.add_scale * __ldg(g.add + out_offset(g, z, m, n, g.add_sz));
      if (g.C_lo) {
        const float h = tf32_round(v);
        g.C[co] = h;
        g.C_lo[co] = tf32_round(v - h);
      } else {
        g.C[co] = v;
      }
    }
  }
}

// ---- skinny variant: N <= 16 (the K-logit head layer, every layer of the toy models) -------------------------------
// The 128x128 tile above wastes > 87 % of its FMAs when N = 10.  Here a CTA owns 64 rows x all (<= 16) columns;
// thread (row r, column group cg) keeps 4 accumulators; A tile [32 k][64 m] and B tile [32 k][16 n] in shared memory.
constexpr int SB_M = 64, SB_K = 32, SB_N = 16;

template <bool A_KC, bool GATHER>
__global__ void __launch_bounds__(NT) gemm_skinny_kernel(DevGemm g) {
  __shared__ float As[SB_K][SB_M + 1];
  __shared__ __align__(16) float Bs[SB_K][SB_N];
  const int z = blockIdx.z;
  const int ks = g.ksplit > 1 ? g.ksplit : 1;
  const int slice = blockIdx.y % ks;
  const int m0 = (blockIdx.y / ks) * SB_M;
  const int tid = threadIdx.x;
  const int r = tid >> 2, cg = tid & 3;
  float acc[4] = {0.f, 0.f, 0.f, 0.f};
  const int npairs = g.A2.ptr ? 2 : 1;
  for (int pair = 0; pair < npairs; ++pair) {
    const DevOperand& A = pair ? g.A2 : g.A1;
    const DevOperand& B = pair ? g.B2 : g.B1;
    const int K = pair ? g.K2 : g.K1;
    const float* Ap = A.ptr + (long long)z * A.sz;
    const float* Bp = B.ptr + (long long)z * B.sz;
    const int nk_all = (K + SB_K - 1) / SB_K;
    const int k_begin = (int)((long long)nk_all * slice / ks) * SB_K;
    const int k_end_t = (int)((long long)nk_all * (slice + 1) / ks) * SB_K;
    const int k_end = k_end_t < K ? k_end_t : K;
    for (int k0 = k_begin; k0 < k_end; k0 += SB_K) {
#pragma unroll
      for (int i = 0; i < (SB_M * SB_K) / NT; ++i) {
        const int idx = tid + NT * i;
        int k, rr;
        if (A_KC) { k = idx % SB_K; rr = idx / SB_K; } else { rr = idx % SB_M; k = idx / SB_M; }
        const int gm = m0 + rr, gk = k0 + k;
        float av = 0.f;
        if (gm < g.M && gk < K) {
          if (GATHER && A.conv.mode) {
            long long off;
            if (conv_offset(A.conv, (uint32_t)gm, (uint32_t)gk, &off)) av = __ldg(Ap + off);
          } else {
            av = __ldg(Ap + (long long)gm * A.s0 + (long long)gk * A.s1);
          }
        }
        As[k][rr] = av;
      }
#pragma unroll
      for (int i = 0; i < (SB_K * SB_N) / NT; ++i) {
        const int idx = tid + NT * i;
        const int n = idx % SB_N, k = idx / SB_N;
        const int gk = k0 + k;
        Bs[k][n] = (n < g.N && gk < K) ? __ldg(Bp + (long long)gk * B.s0 + (long long)n * B.s1) : 0.f;
      }
      __syncthreads();
#pragma unroll
      for (int k = 0; k < SB_K; ++k) {
        const float a = As[k][r];
        const float4 b = *reinterpret_cast<const float4*>(&Bs[k][cg * 4]);
        acc[0] = fmaf(a, b.x, acc[0]); acc[1] = fmaf(a, b.y, acc[1]);
        acc[2] = fmaf(a, b.z, acc[2]); acc[3] = fmaf(a, b.w, acc[3]);
      }
      __syncthreads();
    }
  }
  const int m = m0 + r;
  if (m >= g.M) return;
  if (ks > 1) {
    float* pt = g.part + (long long)slice * g.part_sz + (long long)z * g.M * g.N + (long long)m * g.N;
#pragma unroll
    for (int j = 0; j < 4; ++j) if (cg * 4 + j < g.N) pt[cg * 4 + j] = acc[j];
    return;
  }
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const int n = cg * 4 + j;
    if (n >= g.N) continue;
    float v = g.scale * acc[j];
    if (g.bias) v += __ldg(g.bias + (long long)z * g.bias_sz + n);
    const long long co = (long long)z * g.c_sz + (long long)m * g.c_sm + n;
    if (g.act >= 0) {
      float d;
      v = act_apply(g.act, v, &d);
      if (g.dphi_out) g.dphi_out[co] = d;
    }
    if (g.mask) v *= __ldg(g.mask + (long long)m * g.mask_sm + n);
    if (g.add) v += g.add_scale * __ldg(g.add + (long long)z * g.add_sz + (long long)m * g.c_sm + n);
    if (g.C_lo) {
      const float h = tf32_round(v);
      g.C[co] = h;
      g.C_lo[co] = tf32_round(v - h);
    } else {
      g.C[co] = v;
    }
  }
}

// Row-per-thread form of the skinny GEMM (N <= 16): a thread owns RPT whole output rows (16 accumulators each), so one A element
// from shared memory feeds 16 FMAs and the 16-wide B row is a warp-uniform broadcast (4 LDS.128) - FMA-bound instead of
// shared-memory-bound (the 64 x 16 tile above does 4 FMAs per 2 shared loads).  Tile: RT_K k-steps x (NT * RPT) rows.
constexpr int RT_K = 16;

template <bool A_KC, bool GATHER, int RPT>
__global__ void __launch_bounds__(NT) gemm_rowthread_kernel(DevGemm g) {
  constexpr int TM = NT * RPT;
  __shared__ float As[RT_K][TM + 1];
  __shared__ __align__(16) float Bs[RT_K][SB_N];
  const int z = blockIdx.z;
  const int ks = g.ksplit > 1 ? g.ksplit : 1;
  const int slice = blockIdx.y % ks;
  const int m0 = (blockIdx.y / ks) * TM;
  const int tid = threadIdx.x;
  float acc[RPT][SB_N];
#pragma unroll
  for (int r = 0; r < RPT; ++r)
#pragma unroll
    for (int n = 0; n < SB_N; ++n) acc[r][n] = 0.f;
  const int npairs = g.A2.ptr ? 2 : 1;
  for (int pair = 0; pair < npairs; ++pair) {
    const DevOperand& A = pair ? g.A2 : g.A1;
    const DevOperand& B = pair ? g.B2 : g.B1;
    const int K = pair ? g.K2 : g.K1;
    const float* Ap = A.ptr + (long long)z * A.sz;
    const float* Bp = B.ptr + (long long)z * B.sz;
    const int nk_all = (K + RT_K - 1) / RT_K;
    const int k_begin = (int)((long long)nk_all * slice / ks) * RT_K;
    const int k_end_t = (int)((long long)nk_all * (slice + 1) / ks) * RT_K;
    const int k_end = k_end_t < K ? k_end_t : K;
    for (int k0 = k_begin; k0 < k_end; k0 += RT_K) {
#pragma unroll 4
      for (int i = 0; i < (TM * RT_K) / NT; ++i) {
        const int idx = tid + NT * i;
        int k, rr;
        if (A_KC) { k = idx % RT_K; rr = idx / RT_K; } else { rr = idx % TM; k = idx / TM; }
        const int gm = m0 + rr, gk = k0 + k;
        float av = 0.f;
        if (gm < g.M && gk < K) {
          if (GATHER && A.conv.mode) {
            long long off;
            if (conv_offset(A.conv, (uint32_t)gm, (uint32_t)gk, &off)) av = __ldg(Ap + off);
          } else {
            av = __ldg(Ap + (long long)gm * A.s0 + (long long)gk * A.s1);
          }
        }
        As[k][rr] = av;
      }
      {
        const int n = tid % SB_N, k = tid / SB_N;          // NT == RT_K * SB_N: one element per thread
        const int gk = k0 + k;
        Bs[k][n] = (n < g.N && gk < K) ? __ldg(Bp + (long long)gk * B.s0 + (long long)n * B.s1) : 0.f;
      }
      __syncthreads();
#pragma unroll
      for (int k = 0; k < RT_K; ++k) {
        const float4 b0 = *reinterpret_cast<const float4*>(&Bs[k][0]);
        const float4 b1 = *reinterpret_cast<const float4*>(&Bs[k][4]);
        const float4 b2 = *reinterpret_cast<const float4*>(&Bs[k][8]);
        const float4 b3 = *reinterpret_cast<const float4*>(&Bs[k][12]);
#pragma unroll
        for (int r = 0; r < RPT; ++r) {
          const float a = As[k][tid + r * NT];
          acc[r][0] = fmaf(a, b0.x, acc[r][0]);   acc[r][1] = fmaf(a, b0.y, acc[r][1]);
          acc[r][2] = fmaf(a, b0.z, acc[r][2]);   acc[r][3] = fmaf(a, b0.w, acc[r][3]);
          acc[r][4] = fmaf(a, b1.x, acc[r][4]);   acc[r][5] = fmaf(a, b1.y, acc[r][5]);
          acc[r][6] = fmaf(a, b1.z, acc[r][6]);   acc[r][7] = fmaf(a, b1.w, acc[r][7]);
          acc[r][8] = fmaf(a, b2.x, acc[r][8]);   acc[r][9] = fmaf(a, b2.y, acc[r][9]);
          acc[r][10] = fmaf(a, b2.z, acc[r][10]); acc[r][11] = fmaf(a, b2.w, acc[r][11]);
          acc[r][12] = fmaf(a, b3.x, acc[r][12]); acc[r][13] = fmaf(a, b3.y, acc[r][13]);
          acc[r][14] = fmaf(a, b3.z, acc[r][14]); acc[r][15] = fmaf(a, b3.w, acc[r][15]);
        }
      }
      __syncthreads();
    }
  }
#pragma unroll
  for (int r = 0; r < RPT; ++r) {
    const int m = m0 + tid + r * NT;
    if (m >= g.M) continue;
    if (ks > 1) {
      float* pt = g.part + (long long)slice * g.part_sz + (long long)z * g.M * g.N + (long long)m * g.N;
#pragma unroll
      for (int n = 0; n < SB_N; ++n) if (n < g.N) pt[n] = acc[r][n];
      continue;
    }
#pragma unroll
    for (int n = 0; n < SB_N; ++n) {
      if (n >= g.N) continue;
      float v = g.scale * acc[r][n];
      if (g.bias) v += __ldg(g.bias + (long long)z * g.bias_sz + n);
      const long long co = (long long)z * g.c_sz + (long long)m * g.c_sm + n;
      if (g.act >= 0) {
        float d;
        v = act_apply(g.act, v, &d);
        if (g.dphi_out) g.dphi_out[co] = d;
      }
      if (g.mask) v *= __ldg(g.mask + (long long)m * g.mask_sm + n);
      if (g.add) v += g.add_scale * __ldg(g.add + (long long)z * g.add_sz + (long long)m * g.c_sm + n);
      if (g.C_lo) {
        const float h = tf32_round(v);
        g.C[co] = h;
        g.C_lo[co] = tf32_round(v - h);
      } else {
        g.C[co] = v;
      }
    }
  }
}

// Short-contraction form (K <= 32, both operands k-contiguous, 32 <= N <= 1024): C[m][n] = sum_k A[m][k] B[n][k] is write-bound
// (the delta back-propagation of a narrow conv layer: K = cout = 16, N = kh*kw*cin = 150, 3 GB of patch gradients per call), so the
// whole B sits transposed in shared memory, a thread owns 4 consecutive columns of a row and a warp's stores cover consecutive
// addresses; the 128 x 128 tile kernel reached 0.3 TB/s on it (half-empty second column tile, strided epilogue stores).
constexpr int SK_ROWS = 256;

__global__ void __launch_bounds__(NT) gemm_smallk_kernel(DevGemm g) {
  extern __shared__ __align__(16) float sk_B[];      // [K][Np]
  const int K = g.K1, N = g.N, Np = (N + 3) & ~3;
  const long long z = blockIdx.z;
  const float* Ap = g.A1.ptr + z * g.A1.sz;
  const float* Bp = g.B1.ptr + z * g.B1.sz;
  const int tid = threadIdx.x;
  for (int i = tid; i < K * Np; i += NT) {
    const int k = i / Np, n = i - k * Np;
    sk_B[i] = n < N ? __ldg(Bp + (long long)k * g.B1.s0 + (long long)n * g.B1.s1) : 0.f;
  }
  __syncthreads();
  const int ng = Np >> 2;
  const int rif = NT / ng;                 // rows in flight per block (>= 1: N <= 1024)
  const int cgp = tid % ng, rl = tid / ng;
  if (rl >= rif) return;
  const int n0 = cgp << 2;
  const long long r_end = (long long)(blockIdx.x + 1) * SK_ROWS < g.M ? (long long)(blockIdx.x + 1) * SK_ROWS : g.M;
  float* Cz = g.C + z * g.c_sz;
  const bool pair_ok = ((g.c_sm & 1) == 0) && ((g.c_sz & 1) == 0) && ((reinterpret_cast<uintptr_t>(g.C) & 7) == 0);
  for (long long r = (long long)blockIdx.x * SK_ROWS + rl; r < r_end; r += rif) {
    const float* a = Ap + r * g.A1.s0;
    float acc0 = 0.f, acc1 = 0.f, acc2 = 0.f, acc3 = 0.f;
#pragma unroll 8
    for (int k = 0; k < K; ++k) {
      const float av = __ldg(a + k);
      const float4 b = *reinterpret_cast<const float4*>(&sk_B[k * Np + n0]);
      acc0 = fmaf(av, b.x, acc0); acc1 = fmaf(av, b.y, acc1); acc2 = fmaf(av, b.z, acc2); acc3 = fmaf(av, b.w, acc3);
    }
    float v[4] = {g.scale * acc0, g.scale * acc1, g.scale * acc2, g.scale * acc3};
    if (g.mask) {
#pragma unroll
      for (int j = 0; j < 4; ++j)
        if (n0 + j < N) v[j] *= __ldg(g.mask + r * g.mask_sm + n0 + j);
    }
    float* c = Cz + r * g.c_sm + n0;
    if (pair_ok && n0 + 3 < N) {
      *reinterpret_cast<float2*>(c) = make_float2(v[0], v[1]);
      *reinterpret_cast<float2*>(c + 2) = make_float2(v[2], v[3]);
    } else {
#pragma unroll
      for (int j = 0; j < 4; ++j)
        if (n0 + j < N) c[j] = v[j];
    }
  }
}

// C = epilogue( sum_slices part[slice][z][m][n] ), slices summed in a fixed order
__global__ void splitk_reduce_kernel(DevGemm g, long long total) {
  for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total; idx += (long long)gridDim.x * blockDim.x) {
    const int n = (int)(idx % g.N);
    const long long t = idx / g.N;
    const int m = (int)(t % g.M);
    const long long z = t / g.M;
    float acc = 0.f;
    for (int s = 0; s < g.ksplit; ++s) acc += g.part[(long long)s * g.part_sz + idx];
    float v = g.scale * acc;
    if (g.bias) v += __ldg(g.bias + z * g.bias_sz + n);
    const long long co = out_offset(g, z, m, n, g.c_sz);
    if (g.act >= 0) {
      float d;
      v = act_apply(g.act, v, &d);
      if (g.dphi_out) g.dphi_out[co] = d;
    }
    if (g.mask) v *= __ldg(g.mask + (long long)m * g.mask_sm + n);
    if (g.add) v += g.add_scale * __ldg(g.add + out_offset(g, z, m, n, g.add_sz));
    if (g.C_lo) {
      const float h = tf32_round(v);
      g.C[co] = h;
      g.C_lo[co] = tf32_round(v - h);
    } else {
      g.C[co] = v;
    }
  }
}

}  // namespace

int gemm_simt(const GemmProblem& p_in, cudaStream_t stream) {
  if (p_in.M <= 0 || p_in.N <= 0 || p_in.batch <= 0) return LIP_OK;
  LIP_REQUIRE(p_in.A1.ptr && p_in.B1.ptr && p_in.C, "gemm_simt: null operand");
  // Batch folding: a skinny (N <= 16) batched problem whose A operand is SHARED by the batch entries (the per-probe weight
  // gradients of narrow layers: A = cached activations / patches, B = the probe's deltas) is one wide GEMM
  //   C[m, (z, n)] = sum_k A[m, k] B[z][k, n]
  // with the batch folded into the column index: the 128 x 128 tiles reuse each A element 128 times instead of <= 16 and the
  // shared operand is staged once per 128 / N probes.  Long contractions only (conv weight gradients, K = points x pixels): for the
  // K = 512 head layer of the MNIST MLP the 256 skinny GEMMs were measured 1 % faster per lip_ggn_vp call.  LIP_FOLD_N=0 disables it.
  static const bool fold_on = !(getenv("LIP_FOLD_N") && atoi(getenv("LIP_FOLD_N")) == 0);
  GemmProblem p = p_in;
  int fold_nin = 0;
  long long fold_bsz = 0;
  if (fold_on && p.N <= 16 && p.batch >= 8 && p.A1.sz == 0 && !p.A2.ptr && !p.A1.conv.mode && p.B1.s1 == 1 && !p.epi.bias &&
      !p.epi.mask && p.epi.act < 0 && !p.epi.C_lo && !p.epi.dphi_out && p.batch * p.N < (1LL << 30) && p.K >= 2048) {
    fold_nin = (int)p.N;
    fold_bsz = p.B1.sz;
    p.N = p.batch * p.N;
    p.batch = 1;
  }
  const bool a_kc = p.A1.conv.mode ? (p.A1.conv.mode != 2) : (p.A1.s1 == 1);   // A contiguous along k
  const bool b_kc = (p.B1.s1 != 1);   // B contiguous along k (else along n)
  LIP_REQUIRE(p.A1.conv.mode || a_kc || p.A1.s0 == 1, "gemm_simt: A must be contiguous along m or k");
  LIP_REQUIRE(!b_kc || p.B1.s0 == 1, "gemm_simt: B must be contiguous along k or n");
  if (p.A2.ptr) {
    LIP_REQUIRE(p.B2.ptr != nullptr, "gemm_simt: A2 without B2");
    LIP_REQUIRE(p.A2.conv.mode ? ((p.A2.conv.mode != 2) == a_kc) : (a_kc ? p.A2.s1 == 1 : p.A2.s0 == 1),
                "gemm_simt: A2 must share A1's contiguous index");
    LIP_REQUIRE(b_kc ? p.B2.s0 == 1 : p.B2.s1 == 1, "gemm_simt: B2 must share B1's contiguous index");
  }
  DevGemm g;
  g.M = (int)p.M; g.N = (int)p.N; g.K1 = (int)p.K; g.K2 = (int)p.K2;
  auto cv = [](const GemmOperand& o) {
    DevOperand d{o.ptr, o.sz, o.s0, o.s1, o.conv, 0, FastDiv(), 0};
    if (d.conv.mode) d.conv.finalize();
    return d;
  };
  g.A1 = cv(p.A1); g.B1 = cv(p.B1); g.A2 = cv(p.A2); g.B2 = cv(p.B2);
  g.nin = fold_nin;
  if (fold_nin) {
    g.dnin = FastDiv((uint32_t)fold_nin);
    g.B1.nin = fold_nin; g.B1.dnin = g.dnin; g.B1.fold_sz = fold_bsz;
  }
  g.C = p.C; g.c_sz = p.c_sz; g.c_sm = p.c_sm;
  g.scale = p.epi.scale;
  g.bias = p.epi.bias; g.bias_sz = p.epi.bias_sz;
  g.mask = p.epi.mask; g.mask_sm = p.epi.mask_sm;
  g.add = p.epi.add; g.add_sz = p.epi.add_sz; g.add_scale = p.epi.add_scale;
  g.act = p.epi.act; g.dphi_out = p.epi.dphi_out; g.C_lo = p.epi.C_lo;
  g.ksplit = 1; g.part = nullptr; g.part_sz = 0;
  const bool gather = p.A1.conv.mode != 0 || p.A2.conv.mode != 0;

  if (!p.A2.ptr && !gather && a_kc && b_kc && p.K <= 32 && p.N >= 32 && p.N <= 1024 && !p.epi.bias && p.epi.act < 0 && !p.epi.add &&
      !p.epi.C_lo && !p.epi.dphi_out && !fold_nin && p.batch <= 65535 && p.A1.s1 == 1) {
    const size_t smem = sizeof(float) * (size_t)p.K * (size_t)((p.N + 3) & ~3);
    if (smem > 48 * 1024) LIP_CHECK_CUDA(cudaFuncSetAttribute(gemm_smallk_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024));
    dim3 kg((unsigned)ceil_div(p.M, SK_ROWS), 1, (unsigned)p.batch);
    gemm_smallk_kernel<<<kg, NT, smem, stream>>>(g);
    LIP_LAUNCH_CHECK();
    return LIP_OK;
  }

  // split-K decision: few tiles, long K, scratch available, single operand pair, batch in one launch
  const bool skinny = p.N <= SB_N;
  // skinny form: row-per-thread kernel (LIP_SKINNY_ROWTHREAD=0 restores the 64 x 16 tile kernel); 2 rows per thread for tall problems
  // tall problems only (conv rows = points x pixels): with M = 512 rows per probe (the MNIST-MLP head) the 64-row tiles give 4x more
  // CTAs and were measured ~1 % faster per lip_ggn_vp call
  static const bool rowthread_on = !(getenv("LIP_SKINNY_ROWTHREAD") && atoi(getenv("LIP_SKINNY_ROWTHREAD")) == 0);
  const bool rowthread = rowthread_on && p.M >= 4 * NT;
  const int rpt = 2;
  const int sb_m = rowthread ? NT * rpt : SB_M;
  const int bnt = p.N <= 32 ? 32 : (p.N <= 64 ? 64 : BN);       // narrow tiles for 32- / 64-wide outputs
  const int64_t tiles = skinny ? ceil_div(p.M, sb_m) : ceil_div(p.N, bnt) * ceil_div(p.M, BM);
  const int64_t kstep = skinny ? (rowthread ? RT_K : SB_K) : BK;
  if (p.splitk_ws && !p.A2.ptr && p.batch <= 65535 && tiles * p.batch < 16 * 148 && p.K >= 64 * kstep) {
    int64_t S = ceil_div(32 * 148, tiles * p.batch);       // enough CTAs in flight to hide the latency of the long K stream
    const int64_t smax_k = p.K / (8 * kstep);                       // at least 8 k-tiles per slice
    const int64_t smax_ws = p.splitk_ws_elems / (p.batch * p.M * p.N);
    if (S > smax_k) S = smax_k;
    if (S > smax_ws) S = smax_ws;
    if (S > 1 && tiles * S <= 65535) {
      g.ksplit = (int)S; g.part = p.splitk_ws; g.part_sz = p.batch * p.M * p.N;
    }
  }

  LIP_REQUIRE((skinny ? ceil_div(p.M, sb_m) : ceil_div(p.M, BM)) * g.ksplit <= 65535 && ceil_div(p.N, bnt) <= 2147483647LL,
              "gemm_simt: %lld rows exceed the grid limit of this kernel (tile rows x K slices <= 65535)", (long long)p.M);
  dim3 grid((unsigned)ceil_div(p.N, bnt), (unsigned)(ceil_div(p.M, BM) * g.ksplit), 1);
  // gridDim.z is limited to 65535: chunk the batch.
  const int64_t zmax = 65535;
  for (int64_t z0 = 0; z0 < p.batch; z0 += zmax) {
    DevGemm gz = g;
    int64_t zc = p.batch - z0 < zmax ? p.batch - z0 : zmax;
    gz.A1.ptr += z0 * g.A1.sz; gz.B1.ptr += z0 * g.B1.sz;
    if (gz.A2.ptr) { gz.A2.ptr += z0 * g.A2.sz; gz.B2.ptr += z0 * g.B2.sz; }
    gz.C += z0 * g.c_sz;
    if (gz.bias) gz.bias += z0 * g.bias_sz;
    if (gz.add) gz.add += z0 * g.add_sz;
    if (gz.dphi_out) gz.dphi_out += z0 * g.c_sz;
    if (gz.C_lo) gz.C_lo += z0 * g.c_sz;
    grid.z = (unsigned)zc;
    if (p.N <= SB_N) {
      dim3 sg(1, (unsigned)(ceil_div(p.M, sb_m) * g.ksplit), (unsigned)zc);
      if (rowthread) {
#define LIP_RT_LAUNCH(AK, G)                                                                     \
        do {                                                                                      \
          if (rpt == 2) gemm_rowthread_kernel<AK, G, 2><<<sg, NT, 0, stream>>>(gz);               \
          else gemm_rowthread_kernel<AK, G, 1><<<sg, NT, 0, stream>>>(gz);                        \
        } while (0)
        if (gather) { if (a_kc) LIP_RT_LAUNCH(true, true); else LIP_RT_LAUNCH(false, true); }
        else { if (a_kc) LIP_RT_LAUNCH(true, false); else LIP_RT_LAUNCH(false, false); }
#undef LIP_RT_LAUNCH
      } else if (gather) {
        if (a_kc) gemm_skinny_kernel<true, true><<<sg, NT, 0, stream>>>(gz);
        else gemm_skinny_kernel<false, true><<<sg, NT, 0, stream>>>(gz);
      } else {
        if (a_kc) gemm_skinny_kernel<true, false><<<sg, NT, 0, stream>>>(gz);
        else gemm_skinny_kernel<false, false><<<sg, NT, 0, stream>>>(gz);
      }
      LIP_LAUNCH_CHECK();
      if (g.ksplit > 1) {
        const long long total = (long long)p.batch * p.M * p.N;
        long long rb = (total + 255) / 256;
        if (rb > 148 * 16) rb = 148 * 16;
        splitk_reduce_kernel<<<(unsigned)rb, 256, 0, stream>>>(gz, total);
        LIP_LAUNCH_CHECK();
      }
      continue;
    }
#define LIP_SIMT_LAUNCH_G(AK, BK_, G)                                                            \
    do {                                                                                          \
      if (bnt == 32) gemm_simt_kernel<AK, BK_, 32, G><<<grid, NT, 0, stream>>>(gz);                \
      else if (bnt == 64) gemm_simt_kernel<AK, BK_, 64, G><<<grid, NT, 0, stream>>>(gz);           \
      else gemm_simt_kernel<AK, BK_, 128, G><<<grid, NT, 0, stream>>>(gz);                         \
    } while (0)
#define LIP_SIMT_LAUNCH(AK, BK_)                                                                 \
    do {                                                                                          \
      if (gather) LIP_SIMT_LAUNCH_G(AK, BK_, true); else LIP_SIMT_LAUNCH_G(AK, BK_, false);         \
    } while (0)
    if (a_kc && !b_kc) LIP_SIMT_LAUNCH(true, false);
    else if (!a_kc && !b_kc) LIP_SIMT_LAUNCH(false, false);
    else if (a_kc && b_kc) LIP_SIMT_LAUNCH(true, true);
    else LIP_SIMT_LAUNCH(false, true);
#undef LIP_SIMT_LAUNCH_G
#undef LIP_SIMT_LAUNCH
    LIP_LAUNCH_CHECK();
    if (g.ksplit > 1) {
      const long long total = (long long)p.batch * p.M * p.N;
      long long rb = (total + 255) / 256;
      if (rb > 148 * 16) rb = 148 * 16;
      splitk_reduce_kernel<<<(unsigned)rb, 256, 0, stream>>>(gz, total);
      LIP_LAUNCH_CHECK();
    }
  }
  return LIP_OK;
}

}  // namespace lip
