// lip_krylov.cu — the Krylov recurrences of the hot path as on-stream composite entry points.
//
// Reference semantics (all third-party to the reference, see oracle/lip_oracle.py):
//   matfree.decomp.tridiag_sym(k), reortho="full"      call sites src/sample.py:114, tests/test_sample.py:338
//   matfree.decomp.bidiag(k)  (Golub-Kahan-Lanczos)     call site  src/train_inducing.py:156
//   matfree.funm.funm_lanczos_sym / integrand_funm_sym / integrand_funm_product_logdet
//                                                      src/sample.py:115, src/matfree_monkeypatch.py:25-41, src/train_inducing.py:157
//   jax.scipy.sparse.linalg.cg                          src/stochtrace.py:146,192, src/sample.py:71
//
// Round 1 drove these loops from Python (17 590 launches and ~2 500 ctypes calls for one k = 409 logdet, tensors allocated inside
// the loop, a .item() every 4 CG iterations).  Here ONE C-ABI call enqueues the whole recurrence on the caller's stream: the
// operator is described by a lip_linop (a bound lip_model in one of its roles, a dense symmetric matrix, or a caller callback),
// every scalar (norms, Lanczos / GKL coefficients, CG step sizes) stays on the device, nothing is allocated, and the host never
// waits (CG: optional polling of a pinned flag for early exit).  The vector kernels are fused per recurrence step:
//   axpy_norm   out = s1*x1 | x2  - c*y,  |out| by a last-block reduction           (GKL: u = A v - beta u_prev, alpha = |u|)
//   project     h = Q w: warp-per-row dots against a shared-memory chunk of w, per-CTA partials, last-block fixed-order sum
//   subtract    out = w/s - sum_j (h_j/s) Q_j, |out| by a last-block reduction      (the CGS pass + the norm of its result)
//   scale_store q = w/|w| into the basis row AND the contiguous operand(s) of the next mat-vec
// 8 vector launches per GKL step (round 1: ~30), 6 per Lanczos step with two CGS passes.
// All reductions are deterministic: per-CTA partials summed in a fixed order by the last CTA to arrive (an integer ticket is
// the only atomic), so results do not depend on scheduling.
#include <stdio.h>
#include <stdlib.h>

#include <algorithm>
#include <vector>

#include "lip_comm.cuh"
#include "lip_model.cuh"

using namespace lip;

namespace {

constexpr int VT = 256;
constexpr int RCHUNK = 1024;   // floats of w staged in shared memory per sub-chunk (one 8-deep batch of float4 loads per lane and row)
constexpr int MAXP = 512;      // most per-column partials (CTAs per column)

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

__device__ __forceinline__ float block_sum(float v, float* sm) {
  v = warp_sum(v);
  const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
  __syncthreads();
  if (l == 0) sm[w] = v;
  __syncthreads();
  float t = (threadIdx.x < (blockDim.x >> 5)) ? sm[threadIdx.x] : 0.f;
  if (w == 0) t = warp_sum(t);
  if (threadIdx.x == 0) sm[0] = t;
  __syncthreads();
  t = sm[0];
  __syncthreads();
  return t;
}

// The calling CTA has written its partial results; returns true in exactly one CTA per column: the last to arrive.
__device__ __forceinline__ bool last_block(unsigned* counter, unsigned total) {
  __shared__ int is_last;
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) {
    const unsigned t = atomicAdd(counter, 1u);
    is_last = (t == total - 1);
    if (is_last) *counter = 0;          // ready for the next launch on this stream
  }
  __syncthreads();
  if (is_last) __threadfence();
  return is_last != 0;
}

// sum of part[0..np) in a fixed order; result valid in every thread
__device__ __forceinline__ float sum_fixed(const volatile float* part, int np, float* sm) {
  float v = 0.f;
  for (int i = threadIdx.x; i < np; i += blockDim.x) v += part[i];
  return block_sum(v, sm);
}

inline bool al16(const void* p) { return ((uintptr_t)p & 15) == 0; }

// CTAs per column of the projection kernel: every CTA gets the same number q of RCHUNK sub-chunks (balanced), about 4 x 148 CTAs
// over all columns; fewer, fatter CTAs also keep the last-block coefficient sum short
inline int project_ctas(int64_t n, int64_t B) {
  const int64_t nsub = (n + RCHUNK - 1) / RCHUNK;
  // two CTAs per SM over all columns (2 x 148 CTAs: every SM gets the same number), each CTA walking nsub / np sub-chunks in a
  // strided order: with ~5 B sub-chunks per CTA at n = 1.5 M the imbalance between CTAs is one sub-chunk in 5 B or less
  int64_t np = (2 * 148 + B - 1) / B;
  if (np < 8) np = 8;
  if (np > 512) np = 512;
  if (np > nsub) np = nsub;
  return (int)(np < 1 ? 1 : np);
}

inline int project_ctas_max(int64_t B) {      // upper bound of project_ctas over all n
  int64_t np = (2 * 148 + B - 1) / B;
  if (np < 8) np = 8;
  if (np > 512) np = 512;
  return (int)np;
}

inline int column_ctas(int64_t n, int64_t B, int per_cta) {
  int64_t want = ceil_div(4 * 148, B);
  if (want < 4) want = 4;
  const int64_t most = ceil_div(n, per_cta);
  if (want > most) want = most;
  if (want > MAXP) want = MAXP;
  return (int)(want < 1 ? 1 : want);
}

// ---- axpy_norm -----------------------------------------------------------------------------------------------------
// out[b][j] = X[b][j] - c[b] * y[b][j]     X = s1 * x1[b][j] for j < n1, x2[b][j - n1] for j >= n1
// nrm[b] = sqrt(sum_j out^2)  (optional).  c = coef[b] (NULL: no y term).
struct AxpyNormArgs {
  const float* x1; int64_t ld1; int64_t n1; float s1;
  const float* x2; int64_t ld2;
  const float* y; int64_t ldy; int64_t ysb;   // ysb: batch stride of y
  const float* coef;
  float* out; int64_t ldo;
  float* part; float* nrm; unsigned* counter;
  int64_t n;
  int sq_out;      // 1: nrm receives the sum of squares (a rank-local partial that is all-reduced before its square root)
};

template <int VEC>
__global__ void __launch_bounds__(VT) axpy_norm_kernel(AxpyNormArgs a) {
  __shared__ float sm[32];
  const int b = blockIdx.y, np = gridDim.x;
  const float c = a.coef ? a.coef[b] : 0.f;
  const float* x1 = a.x1 + (int64_t)b * a.ld1;
  const float* x2 = a.x2 ? a.x2 + (int64_t)b * a.ld2 : nullptr;
  const float* y = a.coef ? a.y + (int64_t)b * a.ysb : nullptr;
  float* out = a.out + (int64_t)b * a.ldo;
  float acc = 0.f;
  if (VEC == 4) {
    const int64_t n4 = a.n >> 2, n14 = a.n1 >> 2;    // n1 % 4 == 0 is checked by the launcher in this mode
    for (int64_t i = blockIdx.x * (int64_t)VT + threadIdx.x; i < n4; i += (int64_t)np * VT) {
      float4 v;
      if (i < n14) {
        v = __ldg(reinterpret_cast<const float4*>(x1) + i);
        v.x *= a.s1; v.y *= a.s1; v.z *= a.s1; v.w *= a.s1;
      } else {
        v = __ldg(reinterpret_cast<const float4*>(x2) + (i - n14));
      }
      if (y) {
        const float4 q = __ldg(reinterpret_cast<const float4*>(y) + i);
        v.x -= c * q.x; v.y -= c * q.y; v.z -= c * q.z; v.w -= c * q.w;
      }
      reinterpret_cast<float4*>(out)[i] = v;
      acc += (v.x * v.x + v.y * v.y) + (v.z * v.z + v.w * v.w);
    }
    if (blockIdx.x == 0 && threadIdx.x < (a.n & 3)) {
      const int64_t j = (n4 << 2) + threadIdx.x;
      float v = (j < a.n1) ? a.s1 * x1[j] : x2[j - a.n1];
      if (y) v -= c * y[j];
      out[j] = v;
      acc += v * v;
    }
  } else {
    for (int64_t j = blockIdx.x * (int64_t)VT + threadIdx.x; j < a.n; j += (int64_t)np * VT) {
      float v = (j < a.n1) ? a.s1 * x1[j] : x2[j - a.n1];
      if (y) v -= c * y[j];
      out[j] = v;
      acc += v * v;
    }
  }
  if (!a.nrm) return;
  acc = block_sum(acc, sm);
  if (threadIdx.x == 0) a.part[(int64_t)b * np + blockIdx.x] = acc;
  if (last_block(a.counter + b, np)) {
    const float t = sum_fixed(a.part + (int64_t)b * np, np, sm);
    if (threadIdx.x == 0) a.nrm[b] = a.sq_out ? t : sqrtf(t);
  }
}

// ---- project: h[b][j] = sum_i Q[b][j][i] * w[b][i],  j < kk ----------------------------------------------------------
struct ProjectArgs {
  const float* Q; int64_t ldq; int64_t qsb; int kk;
  const float* w; int64_t ldw;
  float* part; int64_t kpad;     // part[b][cta][kpad]
  float* h;                      // h[b][kpad]
  unsigned* counter;
  int64_t n;
  const int* gate;               // optional: columns with gate[b] == 0 are skipped (conditional second re-orthogonalisation pass)
};

__global__ void __launch_bounds__(VT) project_kernel(ProjectArgs a) {
  extern __shared__ __align__(16) float smem[];
  float* ws = smem;              // RCHUNK
  float* acc = smem + RCHUNK;    // kk
  __shared__ float red[8][33];
  const int b = blockIdx.y, cta = blockIdx.x, np = gridDim.x;
  if (a.gate && a.gate[b] == 0) return;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = VT >> 5;
  const int kk = a.kk;
  for (int j = threadIdx.x; j < kk; j += VT) acc[j] = 0.f;
  const int64_t nsub = (a.n + RCHUNK - 1) / RCHUNK;
  const float* wb = a.w + (int64_t)b * a.ldw;
  const float* Qb = a.Q + (int64_t)b * a.qsb;
  for (int64_t sub = cta; sub < nsub; sub += np) {
    const int64_t i0 = sub * RCHUNK;
    const int len = (int)((a.n - i0) < RCHUNK ? (a.n - i0) : RCHUNK);
    __syncthreads();
    for (int i = threadIdx.x; i < RCHUNK; i += VT) ws[i] = (i < len) ? wb[i0 + i] : 0.f;   // w may be unpadded (ldw = n)
    __syncthreads();
    const int len4 = (len + 3) >> 2;          // rows of Q are zero-padded to ldq (a multiple of 4) and ws is zero beyond len
    const float4* w4 = reinterpret_cast<const float4*>(ws);
    const float* Qs = Qb + i0;
    if (len4 == RCHUNK / 4) {
      // full sub-chunk: a lane holds 8 float4 of the row (one 4 KB row segment per warp) and the NEXT row's 8 loads are issued
      // before the current row is reduced, so every warp keeps 4 - 8 KB in flight (16 warps per SM: enough to cover HBM latency)
      float4 qa[8], qb[8];
      int j = warp;
      if (j < kk) {
        const float4* q4 = reinterpret_cast<const float4*>(Qs + (int64_t)j * a.ldq);
#pragma unroll
        for (int u = 0; u < 8; ++u) qa[u] = __ldg(q4 + lane + 32 * u);
      }
      while (j < kk) {
        const int j1 = j + nw;
        if (j1 < kk) {
          const float4* q4 = reinterpret_cast<const float4*>(Qs + (int64_t)j1 * a.ldq);
#pragma unroll
          for (int u = 0; u < 8; ++u) qb[u] = __ldg(q4 + lane + 32 * u);
        }
        float s = 0.f;
#pragma unroll
        for (int u = 0; u < 8; ++u) {
          const float4 c = w4[lane + 32 * u];
          s += (qa[u].x * c.x + qa[u].y * c.y) + (qa[u].z * c.z + qa[u].w * c.w);
        }
        s = warp_sum(s);
        if (lane == 0) acc[j] += s;             // row j is always handled by the same warp
        if (j1 >= kk) break;
        const int j2 = j1 + nw;
        if (j2 < kk) {
          const float4* q4 = reinterpret_cast<const float4*>(Qs + (int64_t)j2 * a.ldq);
#pragma unroll
          for (int u = 0; u < 8; ++u) qa[u] = __ldg(q4 + lane + 32 * u);
        }
        s = 0.f;
#pragma unroll
        for (int u = 0; u < 8; ++u) {
          const float4 c = w4[lane + 32 * u];
          s += (qb[u].x * c.x + qb[u].y * c.y) + (qb[u].z * c.z + qb[u].w * c.w);
        }
        s = warp_sum(s);
        if (lane == 0) acc[j1] += s;
        j = j2;
      }
    } else {
      for (int j = warp; j < kk; j += nw) {
        const float4* q4 = reinterpret_cast<const float4*>(Qs + (int64_t)j * a.ldq);
        float s = 0.f;
#pragma unroll 4
        for (int i = lane; i < len4; i += 32) {
          const float4 q = __ldg(q4 + i), c = w4[i];
          s += (q.x * c.x + q.y * c.y) + (q.z * c.z + q.w * c.w);
        }
        s = warp_sum(s);
        if (lane == 0) acc[j] += s;
      }
    }
  }
  __syncthreads();
  float* pb = a.part + ((int64_t)b * np + cta) * a.kpad;
  for (int j = threadIdx.x; j < kk; j += VT) pb[j] = acc[j];
  if (!last_block(a.counter + b, np)) return;
  // fixed-order sum over the np CTAs: thread (tx, ty) adds the partials of coefficient j = j0 + 32 t + tx over CTAs c = ty, ty + 8, ...
  // (4 CTAs x 4 tiles = 16 independent L2 loads in flight per thread), then the 8 lanes are added in lane order
  const float* P = a.part + (int64_t)b * np * a.kpad;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  for (int j0 = 0; j0 < kk; j0 += 128) {
    float s[4] = {0.f, 0.f, 0.f, 0.f};
    int c = ty;
    for (; c + 24 < np; c += 32) {
      float v[4][4];
#pragma unroll
      for (int u = 0; u < 4; ++u)
#pragma unroll
        for (int t = 0; t < 4; ++t) {
          const int j = j0 + t * 32 + tx;
          v[u][t] = (j < kk) ? __ldcg(P + (int64_t)(c + 8 * u) * a.kpad + j) : 0.f;
        }
#pragma unroll
      for (int t = 0; t < 4; ++t) s[t] += (v[0][t] + v[1][t]) + (v[2][t] + v[3][t]);
    }
    for (; c < np; c += 8) {
#pragma unroll
      for (int t = 0; t < 4; ++t) {
        const int j = j0 + t * 32 + tx;
        if (j < kk) s[t] += __ldcg(P + (int64_t)c * a.kpad + j);
      }
    }
#pragma unroll
    for (int t = 0; t < 4; ++t) {
      red[ty][tx] = s[t];
      __syncthreads();
      if (ty == 0) {
        float tot = 0.f;
#pragma unroll
        for (int y = 0; y < 8; ++y) tot += red[y][tx];
        const int j = j0 + t * 32 + tx;
        if (j < kk) a.h[(int64_t)b * a.kpad + j] = tot;
      }
      __syncthreads();
    }
  }
}

// ---- subtract: out[b][i] = (w[b][i] - sum_j h[b][j] Q[b][j][i]) * inv,  inv = 1 / scal[b] (scal NULL: 1) ---------------
// nrm[b] = |out| (optional).  In-place (out == w) is fine: every thread reads its own elements before writing them.
struct SubtractArgs {
  const float* Q; int64_t ldq; int64_t qsb; int kk;
  const float* h; int64_t kpad;
  const float* w; int64_t ldw; int w_scalar;      // w_scalar: w rows are not 16-byte aligned (contiguous [B, n] mat-vec output)
  const float* scal;
  float* out; int64_t ldo;
  float* part; float* nrm; unsigned* counter;
  int64_t n;
  int sq_out;
  const int* gate;     // optional: columns with gate[b] == 0 are skipped
};

__global__ void __launch_bounds__(VT) subtract_kernel(SubtractArgs a) {
  extern __shared__ float hs[];
  __shared__ float sm[32];
  const int b = blockIdx.y, np = gridDim.x, kk = a.kk;
  if (a.gate && a.gate[b] == 0) return;
  const float inv = a.scal ? 1.f / a.scal[b] : 1.f;
  for (int j = threadIdx.x; j < kk; j += VT) hs[j] = a.h[(int64_t)b * a.kpad + j];
  __syncthreads();
  const float* Qb = a.Q + (int64_t)b * a.qsb;
  const float* wb = a.w + (int64_t)b * a.ldw;
  float* ob = a.out + (int64_t)b * a.ldo;
  const int64_t n4 = (a.n + 3) >> 2;          // padded rows: the tail lanes of the last float4 are zeros in w and Q
  const int64_t step = a.ldq >> 2;
  float nacc = 0.f;
  for (int64_t i = blockIdx.x * (int64_t)VT + threadIdx.x; i < n4; i += (int64_t)np * VT) {
    float4 acc;
    if (a.w_scalar) {
      const int64_t j0 = i << 2;
      acc.x = wb[j0];
      acc.y = (j0 + 1 < a.n) ? wb[j0 + 1] : 0.f;
      acc.z = (j0 + 2 < a.n) ? wb[j0 + 2] : 0.f;
      acc.w = (j0 + 3 < a.n) ? wb[j0 + 3] : 0.f;
    } else {
      acc = *reinterpret_cast<const float4*>(wb + (i << 2));
    }
    const float4* q = reinterpret_cast<const float4*>(Qb) + i;
    int j = 0;
    for (; j + 4 <= kk; j += 4) {
      const float4 q0 = __ldg(q + (int64_t)(j + 0) * step), q1 = __ldg(q + (int64_t)(j + 1) * step);
      const float4 q2 = __ldg(q + (int64_t)(j + 2) * step), q3 = __ldg(q + (int64_t)(j + 3) * step);
      const float h0 = hs[j], h1 = hs[j + 1], h2 = hs[j + 2], h3 = hs[j + 3];
      acc.x -= h0 * q0.x + h1 * q1.x + h2 * q2.x + h3 * q3.x;
      acc.y -= h0 * q0.y + h1 * q1.y + h2 * q2.y + h3 * q3.y;
      acc.z -= h0 * q0.z + h1 * q1.z + h2 * q2.z + h3 * q3.z;
      acc.w -= h0 * q0.w + h1 * q1.w + h2 * q2.w + h3 * q3.w;
    }
    for (; j < kk; ++j) {
      const float4 q0 = __ldg(q + (int64_t)j * step);
      const float h0 = hs[j];
      acc.x -= h0 * q0.x; acc.y -= h0 * q0.y; acc.z -= h0 * q0.z; acc.w -= h0 * q0.w;
    }
    acc.x *= inv; acc.y *= inv; acc.z *= inv; acc.w *= inv;
    *reinterpret_cast<float4*>(ob + (i << 2)) = acc;
    nacc += (acc.x * acc.x + acc.y * acc.y) + (acc.z * acc.z + acc.w * acc.w);
  }
  if (!a.nrm) return;
  nacc = block_sum(nacc, sm);
  if (threadIdx.x == 0) a.part[(int64_t)b * np + blockIdx.x] = nacc;
  if (last_block(a.counter + b, np)) {
    const float t = sum_fixed(a.part + (int64_t)b * np, np, sm);
    if (threadIdx.x == 0) a.nrm[b] = a.sq_out ? t : sqrtf(t);
  }
}

// ---- short vectors (n <= SHORT_N): many basis rows, few columns ------------------------------------------------------------------
// The kernels above cut the COLUMNS of a long vector over the CTAs; on a short one (the reduced u vectors of the GKL logdet, k + d =
// 5 532 floats at C3b; the d = M K vectors of the sampler's Gram-space Lanczos) that leaves 6 CTAs walking hundreds of rows one after
// the other (60 us per launch at 300 rows).  Here the ROWS carry the parallelism: one warp per basis row for the coefficients, and
// (64 float4 columns) x (4 row groups) per CTA for the update, summed over the row groups in a fixed order.
constexpr int64_t SHORT_N = 16384;
constexpr int SCG = 64, SRG = VT / SCG;

__global__ void __launch_bounds__(VT) project_short_kernel(ProjectArgs a) {
  const int b = blockIdx.y;
  if (a.gate && a.gate[b] == 0) return;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int j = blockIdx.x * (VT / 32) + warp;
  if (j >= a.kk) return;
  const float* wb = a.w + (int64_t)b * a.ldw;
  const float4* q = reinterpret_cast<const float4*>(a.Q + (int64_t)b * a.qsb + (int64_t)j * a.ldq);
  const bool wal = (((uintptr_t)wb) & 15) == 0;
  const int64_t n4 = (a.n + 3) >> 2;
  float acc = 0.f;
#pragma unroll 4
  for (int64_t i = lane; i < n4; i += 32) {
    const float4 qv = __ldg(q + i);
    float4 wv;
    if (wal) {
      wv = __ldg(reinterpret_cast<const float4*>(wb) + i);
    } else {
      const int64_t j0 = i << 2;
      wv.x = wb[j0];
      wv.y = (j0 + 1 < a.n) ? wb[j0 + 1] : 0.f;
      wv.z = (j0 + 2 < a.n) ? wb[j0 + 2] : 0.f;
      wv.w = (j0 + 3 < a.n) ? wb[j0 + 3] : 0.f;
    }
    acc += (qv.x * wv.x + qv.y * wv.y) + (qv.z * wv.z + qv.w * wv.w);
  }
  acc = warp_sum(acc);
  if (lane == 0) a.h[(int64_t)b * a.kpad + j] = acc;
}

__global__ void __launch_bounds__(VT) subtract_short_kernel(SubtractArgs a) {
  extern __shared__ __align__(16) float hs[];      // kk coefficients (padded to 4), then SRG x SCG float4 partial sums
  __shared__ float sm[32];
  const int b = blockIdx.y, np = gridDim.x, kk = a.kk;
  if (a.gate && a.gate[b] == 0) return;
  float4* ps = reinterpret_cast<float4*>(hs + ((kk + 3) & ~3));
  const float inv = a.scal ? 1.f / a.scal[b] : 1.f;
  for (int j = threadIdx.x; j < kk; j += VT) hs[j] = a.h[(int64_t)b * a.kpad + j];
  __syncthreads();
  const int cg = threadIdx.x % SCG, rg = threadIdx.x / SCG;
  const int64_t i = (int64_t)blockIdx.x * SCG + cg;          // float4 column
  const int64_t n4 = (a.n + 3) >> 2, step = a.ldq >> 2;
  float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
  if (i < n4) {
    const float4* q = reinterpret_cast<const float4*>(a.Q + (int64_t)b * a.qsb) + i;
#pragma unroll 8
    for (int j = rg; j < kk; j += SRG) {
      const float4 q0 = __ldg(q + (int64_t)j * step);
      const float h0 = hs[j];
      acc.x = fmaf(h0, q0.x, acc.x); acc.y = fmaf(h0, q0.y, acc.y); acc.z = fmaf(h0, q0.z, acc.z); acc.w = fmaf(h0, q0.w, acc.w);
    }
  }
  ps[rg * SCG + cg] = acc;
  __syncthreads();
  float nacc = 0.f;
  if (rg == 0 && i < n4) {
    const float* wb = a.w + (int64_t)b * a.ldw;
    float4 w;
    if (a.w_scalar) {
      const int64_t j0 = i << 2;
      w.x = wb[j0];
      w.y = (j0 + 1 < a.n) ? wb[j0 + 1] : 0.f;
      w.z = (j0 + 2 < a.n) ? wb[j0 + 2] : 0.f;
      w.w = (j0 + 3 < a.n) ? wb[j0 + 3] : 0.f;
    } else {
      w = *reinterpret_cast<const float4*>(wb + (i << 2));
    }
    float4 t = ps[cg];
#pragma unroll
    for (int g = 1; g < SRG; ++g) {
      const float4 u = ps[g * SCG + cg];
      t.x += u.x; t.y += u.y; t.z += u.z; t.w += u.w;
    }
    w.x = (w.x - t.x) * inv; w.y = (w.y - t.y) * inv; w.z = (w.z - t.z) * inv; w.w = (w.w - t.w) * inv;
    *reinterpret_cast<float4*>(a.out + (int64_t)b * a.ldo + (i << 2)) = w;
    nacc = (w.x * w.x + w.y * w.y) + (w.z * w.z + w.w * w.w);
  }
  if (!a.nrm) return;
  nacc = block_sum(nacc, sm);
  if (threadIdx.x == 0) a.part[(int64_t)b * np + blockIdx.x] = nacc;
  if (last_block(a.counter + b, np)) {
    const float t = sum_fixed(a.part + (int64_t)b * np, np, sm);
    if (threadIdx.x == 0) a.nrm[b] = a.sq_out ? t : sqrtf(t);
  }
}

// ---- scale_store: y = x / scal[b] into up to three destinations -------------------------------------------------------
//   o1 (stride ld1, optional): all n columns              (the basis row)
//   o2 (stride ld2, optional): columns [0, n2)            (contiguous operand of the next mat-vec)
//   o3 (stride ld3, optional): columns [n2, n)            (second contiguous operand: the output-space part of a GKL u vector)
struct ScaleStoreArgs {
  const float* x; int64_t ldx;
  const float* scal;
  float* o1; int64_t ld1; int64_t o1sb;
  float* o2; int64_t ld2; int64_t n2;
  float* o3; int64_t ld3;
  int64_t n;
};

__global__ void __launch_bounds__(VT) scale_store_kernel(ScaleStoreArgs a) {
  const int b = blockIdx.y;
  const float inv = a.scal ? 1.f / a.scal[b] : 1.f;
  const float* xb = a.x + (int64_t)b * a.ldx;
  float* o1 = a.o1 ? a.o1 + (int64_t)b * a.o1sb : nullptr;
  float* o2 = a.o2 ? a.o2 + (int64_t)b * a.ld2 : nullptr;
  float* o3 = a.o3 ? a.o3 + (int64_t)b * a.ld3 : nullptr;
  for (int64_t j = blockIdx.x * (int64_t)VT + threadIdx.x; j < a.n; j += (int64_t)gridDim.x * VT) {
    const float v = xb[j] * inv;
    if (o1) o1[j] = v;
    if (j < a.n2) { if (o2) o2[j] = v; }
    else if (o3) o3[j - a.n2] = v;
  }
}

// ---- tiny bookkeeping kernels ----------------------------------------------------------------------------------------
// x[b] = sqrt(x[b])   (after the all-reduce of rank-local sums of squares)
__global__ void sqrt_kernel(float* x, int B) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b < B) x[b] = sqrtf(x[b]);
}
// dst[b*stride + idx] = src[b]
__global__ void record_kernel(float* dst, int64_t stride, int64_t idx, const float* src, int B) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b < B) dst[(int64_t)b * stride + idx] = src[b];
}
// Lanczos step i: diag[i] = h[i];  off[i-1] = 0.5 * (h[i-1] + len)   (T = (H + H^T)/2 of the Arnoldi form)
// (runs between the first projection of step i and its subtract, so `len` still holds |w| of step i-1 = H[i][i-1])
__global__ void lanczos_record_kernel(float* diag, float* off, const float* h, int64_t kpad, const float* len, int i, int k, int B) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B) return;
  diag[(int64_t)b * k + i] = h[(int64_t)b * kpad + i];
  if (i > 0) off[(int64_t)b * (k - 1) + i - 1] = 0.5f * (h[(int64_t)b * kpad + i - 1] + len[b]);
}
// e[b][i] = val, e[b][i-1] = 0: the coefficient vector of basis vector i (reduced GKL u vectors)
__global__ void unit_kernel(float* e, int64_t ld, int64_t i, float val, int B) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B) return;
  e[(int64_t)b * ld + i] = val;
  if (i > 0) e[(int64_t)b * ld + i - 1] = 0.f;
}
// beta[b] *= nv[b];  dst[b*stride + idx] = beta[b]   (reduced GKL: the norm AFTER the projection is the bidiagonal entry)
// Breakdown: when the projection leaves less than 1e-6 of the vector (the fp32 noise floor is ~1e-7) the Krylov space is exhausted and
// beta_{i+1} = 0 in exact arithmetic.  The column is marked dead: beta_{i+1} and every later beta are recorded as 0, later alphas as 1,
// so the trailing block of the bidiagonal is exactly decoupled from e1 (it carries weight ~1e-13 in the explicit recurrence, which keeps
// iterating on rounding noise there); nv is set to 1 so that the remainder is stored unscaled instead of blown up to unit norm.
__global__ void beta_fix_kernel(float* dst, int64_t stride, int64_t idx, float* beta, float* nv, int* dead, int B) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B) return;
  if (dead[b] || !(nv[b] >= 1e-6f)) {
    dead[b] = 1;
    beta[b] = 0.f;
    nv[b] = 1.f;
    dst[(int64_t)b * stride + idx] = 0.f;
    return;
  }
  const float t = beta[b] * nv[b];
  beta[b] = t;
  dst[(int64_t)b * stride + idx] = t;
}
// dst[b*stride + idx] = dead[b] ? 1 : src[b]
__global__ void record_alive_kernel(float* dst, int64_t stride, int64_t idx, const float* src, const int* dead, int B) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b < B) dst[(int64_t)b * stride + idx] = dead[b] ? 1.f : src[b];
}
// "twice is enough": gate[b] = 1 when the projection removed more than half of the (unit) vector, i.e. cancellation left the
// remainder with a relative error that a second pass has to clean up;  count[0] += number of gated columns (diagnostics)
__global__ void reorth_gate_kernel(int* gate, const float* nv, float thr, int B, int* count) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B) return;
  const int g = !(nv[b] >= thr) ? 1 : 0;       // NaN gates too
  gate[b] = g;
  if (g && count) atomicAdd(count, 1);
}
// gate[b] = 1 when the pass left less than `thr` of what it started from (after[b] < thr * before[b])
__global__ void ratio_gate_kernel(int* gate, const float* after, const float* before, float thr, int B) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b < B) gate[b] = !(after[b] >= thr * before[b]) ? 1 : 0;
}
// nv[b] = gate[b] ? nv2[b] : nv[b]
__global__ void gate_merge_kernel(float* nv, const float* nv2, const int* gate, int B) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b < B && gate[b]) nv[b] = nv2[b];
}
// out[b] = q[b] * nrm[b]^2
__global__ void quad_scale_kernel(float* out, const float* q, const float* nrm, int B) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b < B) out[b] = q[b] * nrm[b] * nrm[b];
}
// c[b][j] *= s[b]
__global__ void rowscale_kernel(float* c, int64_t ld, const float* s, int k, int B) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx < B * k) c[(int64_t)(idx / k) * ld + idx % k] *= s[idx / k];
}
// any[0] = OR_b active[b]
__global__ void any_active_kernel(const int* active, int B, int* any) {
  __shared__ int s;
  if (threadIdx.x == 0) s = 0;
  __syncthreads();
  for (int b = threadIdx.x; b < B; b += blockDim.x) if (active[b]) s = 1;
  __syncthreads();
  if (threadIdx.x == 0) *any = s;
}

// ---- dense symmetric operator  out = alpha * u + beta * (u G),  G [n, n] symmetric row-major, u [B, n] ------------------
// (src/sample.py:120-125: the d x d Gram mat-vec inside the sampler's Lanczos.)  HBM / L2 bound: G is read once per 8 probes.
// Grid (column tiles of 512, row splits); a thread owns 4 consecutive columns (one 16-byte load per row of G) for NB probes;
// the partial sums over its row slice land in part[split][b][j]; the second kernel adds them in split order (deterministic)
// and applies alpha, beta.  n % 4 == 0 and a 16-byte aligned G are required by the vector loads (else the scalar variant).
constexpr int DS_MAXSPLIT = 32;
template <int NB, int VEC>
__global__ void __launch_bounds__(128) dense_sym_partial_kernel(const float* __restrict__ G, int64_t n, const float* __restrict__ u,
                                                                int64_t ldu, int64_t b0, int64_t B, float* __restrict__ part,
                                                                int64_t rows_per_split) {
  __shared__ float us[NB][128];
  const int64_t j = ((int64_t)blockIdx.x * 128 + threadIdx.x) * VEC;
  const int64_t r0 = (int64_t)blockIdx.y * rows_per_split;
  const int64_t r1 = (r0 + rows_per_split < n) ? r0 + rows_per_split : n;
  float acc[NB][VEC];
#pragma unroll
  for (int p = 0; p < NB; ++p)
#pragma unroll
    for (int c = 0; c < VEC; ++c) acc[p][c] = 0.f;
  for (int64_t i0 = r0; i0 < r1; i0 += 128) {
    __syncthreads();
#pragma unroll
    for (int p = 0; p < NB; ++p) {
      const int64_t i = i0 + threadIdx.x;
      us[p][threadIdx.x] = (i < r1 && b0 + p < B) ? u[(b0 + p) * ldu + i] : 0.f;
    }
    __syncthreads();
    if (j < n) {
      const int lim = (int)((r1 - i0) < 128 ? (r1 - i0) : 128);
#pragma unroll 4
      for (int i = 0; i < lim; ++i) {
        float g[VEC];
        if (VEC == 4) {
          const float4 t = __ldg(reinterpret_cast<const float4*>(G + (i0 + i) * n + j));
          g[0] = t.x; g[1] = t.y; g[2] = t.z; g[3] = t.w;
        } else {
          g[0] = __ldg(G + (i0 + i) * n + j);
        }
#pragma unroll
        for (int p = 0; p < NB; ++p) {
          const float uv = us[p][i];
#pragma unroll
          for (int c = 0; c < VEC; ++c) acc[p][c] += uv * g[c];
        }
      }
    }
  }
  if (j < n) {
#pragma unroll
    for (int p = 0; p < NB; ++p)
      if (b0 + p < B) {
#pragma unroll
        for (int c = 0; c < VEC; ++c) part[((int64_t)blockIdx.y * B + b0 + p) * n + j + c] = acc[p][c];
      }
  }
}

__global__ void dense_sym_finish_kernel(const float* __restrict__ part, int nsplit, const float* __restrict__ u, int64_t ldu,
                                        float* __restrict__ out, int64_t ldo, int64_t n, int64_t B, float alpha, float beta) {
  const int64_t idx = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (idx >= n * B) return;
  const int64_t b = idx / n, j = idx - b * n;
  float s = 0.f;
  for (int sp = 0; sp < nsplit; ++sp) s += part[((int64_t)sp * B + b) * n + j];
  out[b * ldo + j] = alpha * u[b * ldu + j] + beta * s;
}


// ---- tall-skinny building blocks of Hutch++ (src/stochtrace.py:118-135): the [n, s1] QR and the deflation G - (G Q) Q^T ------------
// Row-stored blocks: Y [s, n] holds s vectors of length n (what the probe-batched operators produce).
// gram:  P[a][b] = sum_j A[a][j] * Bm[b][j]   accumulated in float64 (the products of floats are exact in double, so the Gram of an
// ill-conditioned block keeps 1e-16 relative accuracy and its Cholesky factor exists for cond(Y) up to ~1e7).  Grid (pair tiles of
// 16 x 16, n splits); deterministic second stage.
constexpr int GT = 16;          // pair tile
constexpr int GC = 64;          // columns per shared-memory tile
__global__ void __launch_bounds__(GT * GT) gram_partial_kernel(const float* __restrict__ A, int64_t lda, int sa, const float* __restrict__ Bm,
                                                               int64_t ldb, int sb, int64_t n, int64_t cols_per_split,
                                                               double* __restrict__ part) {
  __shared__ float As[GT][GC + 1], Bs[GT][GC + 1];
  const int tiles_b = (sb + GT - 1) / GT;
  const int ta = blockIdx.x / tiles_b, tb = blockIdx.x % tiles_b;
  const int ia = threadIdx.x / GT, ib = threadIdx.x % GT;
  const int64_t c0 = (int64_t)blockIdx.y * cols_per_split;
  const int64_t c1 = (c0 + cols_per_split < n) ? c0 + cols_per_split : n;
  double acc = 0.0;
  for (int64_t c = c0; c < c1; c += GC) {
    __syncthreads();
    for (int idx = threadIdx.x; idx < GT * GC; idx += GT * GT) {
      const int r = idx / GC, cc = idx % GC;
      const int64_t col = c + cc;
      const int ra = ta * GT + r, rb = tb * GT + r;
      As[r][cc] = (ra < sa && col < c1) ? A[(int64_t)ra * lda + col] : 0.f;
      Bs[r][cc] = (rb < sb && col < c1) ? Bm[(int64_t)rb * ldb + col] : 0.f;
    }
    __syncthreads();
#pragma unroll 8
    for (int cc = 0; cc < GC; ++cc) acc += (double)As[ia][cc] * (double)Bs[ib][cc];
  }
  const int a = ta * GT + ia, b = tb * GT + ib;
  if (a < sa && b < sb) part[((int64_t)blockIdx.y * sa + a) * sb + b] = acc;
}

// out[a][b] = sum over splits (fixed order); optional float copy
__global__ void gram_finish_kernel(const double* __restrict__ part, int nsplit, int64_t count, double* __restrict__ out, float* __restrict__ outf) {
  const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i >= count) return;
  double s = 0.0;
  for (int sp = 0; sp < nsplit; ++sp) s += part[(int64_t)sp * count + i];
  if (out) out[i] = s;
  if (outf) outf[i] = (float)s;
}

// One CTA: C (s x s, float64, symmetric positive definite up to rounding) -> Tinv = L^{-1} with C + shift I = L L^T, as float [s][s]
// (lower triangular, zeros above).  shift = 1e-13 * trace(C) keeps the factorisation alive on numerically rank-deficient blocks; the
// second and third passes of CholeskyQR remove its effect.  flag[0] is set when a pivot is not positive.
__global__ void __launch_bounds__(1024) chol_inv_kernel(double* __restrict__ C, int s, float* __restrict__ Tinv, double* __restrict__ Linv,
                                                        int* __restrict__ flag) {
  __shared__ double sh_piv;
  __shared__ double red[32];
  const int t = threadIdx.x, nt = blockDim.x;
  double tr = 0.0;
  for (int i = t; i < s; i += nt) tr += C[(int64_t)i * s + i];
  for (int o = 16; o > 0; o >>= 1) tr += __shfl_xor_sync(0xffffffffu, tr, o);
  if ((t & 31) == 0) red[t >> 5] = tr;
  __syncthreads();
  if (t == 0) {
    double tot = 0.0;
    for (int w = 0; w < (nt + 31) / 32; ++w) tot += red[w];
    sh_piv = 1e-13 * tot;
  }
  __syncthreads();
  const double shift = sh_piv;
  __syncthreads();
  for (int j = 0; j < s; ++j) {
    if (t == 0) {
      const double dj = C[(int64_t)j * s + j] + shift;
      if (!(dj > 0.0)) *flag = 1;
      sh_piv = sqrt(dj > 0.0 ? dj : 1.0);
      C[(int64_t)j * s + j] = sh_piv;
    }
    __syncthreads();
    const double piv = sh_piv;
    for (int i = j + 1 + t; i < s; i += nt) C[(int64_t)i * s + j] /= piv;
    __syncthreads();
    // trailing update of the lower triangle: C[i][k] -= L[i][j] L[k][j], j < k <= i
    const int m = s - j - 1;
    for (int64_t idx = t; idx < (int64_t)m * m; idx += nt) {
      const int i = j + 1 + (int)(idx / m), k = j + 1 + (int)(idx % m);
      if (k <= i) C[(int64_t)i * s + k] -= C[(int64_t)i * s + j] * C[(int64_t)k * s + j];
    }
    __syncthreads();
  }
  // columns of L^{-1} by forward substitution, one column per thread
  for (int c = t; c < s; c += nt) {
    for (int i = 0; i < s; ++i) {
      double x = 0.0;
      if (i >= c) {
        x = (i == c) ? 1.0 : 0.0;
        for (int j = c; j < i; ++j) x -= C[(int64_t)i * s + j] * Linv[(int64_t)j * s + c];
        x /= C[(int64_t)i * s + i];
      }
      Linv[(int64_t)i * s + c] = x;
      Tinv[(int64_t)i * s + c] = (float)x;
    }
  }
}

// acc[0] (+)= scale * sum_{i<s} sum_j A[i][j] * Bm[i][j]   (float64 two-stage; the trace terms tr(Q^T X Q), tr(G_perp X G_perp^T))
__global__ void __launch_bounds__(VT) rowdot_partial_kernel(const float* __restrict__ A, int64_t lda, const float* __restrict__ Bm,
                                                            int64_t ldb, int s, int64_t n, double* __restrict__ part) {
  __shared__ double sm[VT / 32];
  double acc = 0.0;
  const int64_t total = (int64_t)s * n;
  for (int64_t idx = blockIdx.x * (int64_t)VT + threadIdx.x; idx < total; idx += (int64_t)gridDim.x * VT) {
    const int64_t i = idx / n, j = idx - i * n;
    acc += (double)A[i * lda + j] * (double)Bm[i * ldb + j];
  }
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  if ((threadIdx.x & 31) == 0) sm[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0.0;
    for (int w = 0; w < VT / 32; ++w) t += sm[w];
    part[blockIdx.x] = t;
  }
}
__global__ void eye_kernel(float* __restrict__ out, int64_t n) {
  const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i < n * n) out[i] = (i / n == i % n) ? 1.f : 0.f;
}
__global__ void rowdot_finish_kernel(const double* __restrict__ part, int np, double scale, double* __restrict__ acc, int accumulate,
                                     float* __restrict__ outf) {
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  double t = 0.0;
  for (int i = 0; i < np; ++i) t += part[i];
  const double v = (accumulate ? acc[0] : 0.0) + scale * t;
  acc[0] = v;
  if (outf) outf[0] = (float)v;
}

// =====================================================================================================================
// host side
// =====================================================================================================================
struct Bump {          // carves the caller's workspace
  char* p; char* end; bool ok = true;
  Bump(void* ws, size_t bytes) : p((char*)align_up((uintptr_t)ws, 256)), end((char*)ws + bytes) {}
  template <class T>
  T* take(size_t count) {
    char* q = p;
    p += align_up(count * sizeof(T), 256);
    if (p > end) ok = false;
    return (T*)q;
  }
};
inline size_t rsz(size_t count, size_t elt) { return align_up(count * elt, 256); }

inline int64_t pad4(int64_t n) { return (n + 3) / 4 * 4; }

struct Op {
  const lip_linop* op;
  int64_t n_in = 0, n_out = 0;   // columns / rows of the operator
  int64_t D = 0, d = 0;          // model kinds
  bool symmetric = false;
  void* mws = nullptr; size_t mws_bytes = 0;   // model workspace
  float* dpart = nullptr; int nsplit = 0;      // dense-sym partials
};

constexpr int64_t MAX_COLUMNS = 65535;     // blockIdx.y of the vector kernels

int op_init(Op& o, const lip_linop* op) {
  LIP_REQUIRE(op, "null operator");
  o.op = op;
  switch (op->kind) {
    case LIP_LINOP_GGN:
      LIP_REQUIRE(op->model && op->model->bound, "lip_linop GGN: model not bound");
      o.D = op->model->D; o.d = op->model->M * op->model->K;
      o.n_in = o.n_out = o.D; o.symmetric = true;
      break;
    case LIP_LINOP_GKL:
      LIP_REQUIRE(op->model && op->model->bound, "lip_linop GKL: model not bound");
      LIP_REQUIRE(op->alpha >= 0.f, "lip_linop GKL: alpha must be >= 0");
      o.D = op->model->D; o.d = op->model->M * op->model->K;
      o.n_in = o.D; o.n_out = o.D + o.d;
      break;
    case LIP_LINOP_DENSE_SYM:
      LIP_REQUIRE(op->dense && op->n > 0, "lip_linop DENSE_SYM: null matrix");
      o.n_in = o.n_out = op->n; o.symmetric = true;
      break;
    case LIP_LINOP_CALLBACK:
      LIP_REQUIRE(op->fn && op->n > 0 && op->n_out > 0 && op->cb_in && op->cb_out, "lip_linop CALLBACK: fn, n, n_out, cb_in, cb_out required");
      o.n_in = op->n; o.n_out = op->n_out; o.symmetric = op->symmetric != 0;
      LIP_REQUIRE(!o.symmetric || o.n_in == o.n_out, "lip_linop CALLBACK: a symmetric operator must be square");
      break;
    default:
      LIP_REQUIRE(false, "lip_linop: unknown kind %d", op->kind);
  }
  return LIP_OK;
}

size_t op_ws_bytes(const Op& o, int64_t B) {
  const lip_linop* op = o.op;
  if (op->kind == LIP_LINOP_GGN || op->kind == LIP_LINOP_GKL) return align_up(lip_workspace_bytes(op->model, B), 256) + 256;
  if (op->kind == LIP_LINOP_DENSE_SYM) return rsz((size_t)DS_MAXSPLIT * B * op->n, 4) + 256;
  return 256;
}

int op_carve(Op& o, Bump& bp, int64_t B) {
  const lip_linop* op = o.op;
  if (op->kind == LIP_LINOP_GGN || op->kind == LIP_LINOP_GKL) {
    o.mws_bytes = lip_workspace_bytes(op->model, B);
    o.mws = bp.take<char>(o.mws_bytes);
  } else if (op->kind == LIP_LINOP_DENSE_SYM) {
    o.dpart = bp.take<float>((size_t)DS_MAXSPLIT * B * op->n);
  }
  return LIP_OK;
}

// contiguous in [B, n_in] -> contiguous out [B, n_out]; symmetric kinds and CALLBACK only (GKL has its own two-segment path)
int op_apply(Op& o, const float* in, float* out, int64_t B, int transpose, cudaStream_t st) {
  const lip_linop* op = o.op;
  if (op->kind == LIP_LINOP_GGN) {
    return lip_ggn_vp(op->model, in, out, B, op->scale, op->alpha, o.mws, o.mws_bytes, st);
  }
  if (op->kind == LIP_LINOP_DENSE_SYM) {
    const int64_t n = op->n;
    const bool v4 = (n % 4 == 0) && al16(op->dense);
    const int64_t tiles = ceil_div(n, v4 ? 512 : 128);
    int nsplit = (int)std::min<int64_t>(DS_MAXSPLIT, std::max<int64_t>(1, (4 * 148) / tiles));
    const int64_t rows = ceil_div(ceil_div(n, nsplit), 32) * 32;
    nsplit = (int)ceil_div(n, rows);
    dim3 grid((unsigned)tiles, (unsigned)nsplit);
    for (int64_t b0 = 0; b0 < B; b0 += 8) {
      if (v4) dense_sym_partial_kernel<8, 4><<<grid, 128, 0, st>>>(op->dense, n, in, n, b0, B, o.dpart, rows);
      else dense_sym_partial_kernel<8, 1><<<grid, 128, 0, st>>>(op->dense, n, in, n, b0, B, o.dpart, rows);
      LIP_LAUNCH_CHECK();
    }
    dense_sym_finish_kernel<<<(unsigned)ceil_div(n * B, 256), 256, 0, st>>>(o.dpart, nsplit, in, n, out, n, n, B, op->alpha,
                                                                          op->beta);
    LIP_LAUNCH_CHECK();
    return LIP_OK;
  }
  if (op->kind == LIP_LINOP_CALLBACK) {
    // the callback reads cb_in / cb_in_t and writes cb_out / cb_out_t (caller-owned, fixed for the whole recurrence)
    const bool tr = transpose && !o.symmetric;
    float* cin = tr ? op->cb_in_t : op->cb_in;
    float* cout = tr ? op->cb_out_t : op->cb_out;
    const int64_t ni = tr ? o.n_out : o.n_in, no = tr ? o.n_in : o.n_out;
    LIP_REQUIRE(cin && cout, "lip_linop CALLBACK: transpose buffers missing");
    if (in != cin) LIP_CHECK_CUDA(cudaMemcpyAsync(cin, in, sizeof(float) * (size_t)B * ni, cudaMemcpyDeviceToDevice, st));
    const int rc = op->fn(op->ctx, tr ? 1 : 0, B, (lip_stream_t)st);
    if (rc != 0) { set_error("lip_linop CALLBACK: the mat-vec callback failed (%d)", rc); return LIP_ERR_INVALID; }
    if (out != cout) LIP_CHECK_CUDA(cudaMemcpyAsync(out, cout, sizeof(float) * (size_t)B * no, cudaMemcpyDeviceToDevice, st));
    return LIP_OK;
  }
  LIP_REQUIRE(false, "op_apply: kind %d has no contiguous form", op->kind);
}

// ---- launch helpers --------------------------------------------------------------------------------------------------
struct Red {            // reduction scratch shared by all kernels of a recurrence (stream order serialises its use)
  float* part = nullptr;        // [B][MAXP]
  unsigned* counter = nullptr;  // [B]
  float* ppart = nullptr;       // projection partials [B][np][kpad]
  float* h = nullptr;           // [B][kpad]
  int64_t kpad = 0;
  int sq_out = 0;               // sharded recurrences: norms leave the kernels as rank-local sums of squares
};

bool short_vector_kernels() {
  static const int on = [] {
    const char* e = getenv("LIP_KRYLOV_SHORT");
    return (e && e[0] == '0') ? 0 : 1;
  }();
  return on != 0;
}

int launch_axpy_norm(AxpyNormArgs a, int64_t B, const Red& r, cudaStream_t st) {
  a.part = r.part; a.counter = r.counter; a.sq_out = r.sq_out;
  const bool v4 = al16(a.x1) && (a.ld1 % 4 == 0) && (a.n1 % 4 == 0 || a.n1 >= a.n) && (!a.x2 || (al16(a.x2) && a.ld2 % 4 == 0)) &&
                  (!a.coef || (al16(a.y) && a.ldy % 4 == 0 && a.ysb % 4 == 0)) && al16(a.out) && (a.ldo % 4 == 0);
  const int np = column_ctas(a.n, B, VT * 4 * 2);
  dim3 grid(np, (unsigned)B);
  if (v4) axpy_norm_kernel<4><<<grid, VT, 0, st>>>(a);
  else axpy_norm_kernel<1><<<grid, VT, 0, st>>>(a);
  LIP_LAUNCH_CHECK();
  return LIP_OK;
}

int launch_project(const float* Q, int64_t ldq, int64_t qsb, int kk, const float* w, int64_t ldw, int64_t n, int64_t B,
                   const Red& r, cudaStream_t st, const int* gate = nullptr) {
  if (kk <= 0) return LIP_OK;
  ProjectArgs a{Q, ldq, qsb, kk, w, ldw, r.ppart, r.kpad, r.h, r.counter, n, gate};
  if (n <= SHORT_N && short_vector_kernels()) {
    project_short_kernel<<<dim3((unsigned)ceil_div(kk, VT / 32), (unsigned)B), VT, 0, st>>>(a);
    LIP_LAUNCH_CHECK();
    return LIP_OK;
  }
  const int np = project_ctas(n, B);
  dim3 grid(np, (unsigned)B);
  project_kernel<<<grid, VT, sizeof(float) * (RCHUNK + (size_t)kk), st>>>(a);
  LIP_LAUNCH_CHECK();
  return LIP_OK;
}

int launch_subtract(const float* Q, int64_t ldq, int64_t qsb, int kk, const float* w, int64_t ldw, const float* scal, float* out,
                    int64_t ldo, float* nrm, int64_t n, int64_t B, const Red& r, cudaStream_t st, const int* gate = nullptr) {
  SubtractArgs a{Q, ldq, qsb, kk, r.h, r.kpad, w, ldw, (!al16(w) || ldw % 4 != 0) ? 1 : 0, scal, out, ldo, r.part, nrm, r.counter, n, r.sq_out,
                 gate};
  if (n <= SHORT_N && short_vector_kernels() && kk <= 4096) {
    const int64_t n4 = (n + 3) >> 2;
    const size_t smem = sizeof(float) * (size_t)(((kk + 3) & ~3) + VT * 4);
    subtract_short_kernel<<<dim3((unsigned)ceil_div(n4, SCG), (unsigned)B), VT, smem, st>>>(a);
    LIP_LAUNCH_CHECK();
    return LIP_OK;
  }
  const int np = column_ctas(n, B, VT * 4);
  dim3 grid(np, (unsigned)B);
  subtract_kernel<<<grid, VT, sizeof(float) * (size_t)(kk > 0 ? kk : 1), st>>>(a);
  LIP_LAUNCH_CHECK();
  return LIP_OK;
}

int launch_scale_store(ScaleStoreArgs a, int64_t B, cudaStream_t st) {
  int64_t g = ceil_div(a.n, VT * 4);
  const int64_t cap = std::max<int64_t>(4, ceil_div(8 * 148, B));
  if (g > cap) g = cap;
  dim3 grid((unsigned)(g < 1 ? 1 : g), (unsigned)B);
  scale_store_kernel<<<grid, VT, 0, st>>>(a);
  LIP_LAUNCH_CHECK();
  return LIP_OK;
}

int set_kernel_limits(int64_t k) {
  const size_t need = sizeof(float) * (RCHUNK + (size_t)k);
  LIP_REQUIRE(need <= 200 * 1024, "Krylov depth %lld too large for the shared-memory coefficient buffers", (long long)k);
  if (need > 48 * 1024) {
    LIP_CHECK_CUDA(cudaFuncSetAttribute(project_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)need));
    LIP_CHECK_CUDA(cudaFuncSetAttribute(subtract_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)need));
  }
  return LIP_OK;
}

size_t red_bytes(int64_t n, int64_t B, int64_t k) {
  const int64_t kpad = pad4(k);
  const int np = project_ctas_max(B);
  return rsz((size_t)B * MAXP, 4) + rsz((size_t)B, 4) + rsz((size_t)B * np * kpad, 4) + rsz((size_t)B * kpad, 4);
}

int red_carve(Red& r, Bump& bp, int64_t n, int64_t B, int64_t k, cudaStream_t st) {
  r.kpad = pad4(k);
  const int np = project_ctas_max(B);
  r.part = bp.take<float>((size_t)B * MAXP);
  r.counter = bp.take<unsigned>((size_t)B);
  r.ppart = bp.take<float>((size_t)B * np * r.kpad);
  r.h = bp.take<float>((size_t)B * r.kpad);
  if (!bp.ok) return LIP_OK;      // the caller reports the workspace error
  LIP_CHECK_CUDA(cudaMemsetAsync(r.counter, 0, sizeof(unsigned) * (size_t)B, st));
  return LIP_OK;
}

// ---- D-sharding of the Krylov vectors over the ranks of a lip_comm ---------------------------------------------------------
// Rank r of S owns columns [off, off + nloc) of every parameter-space vector (and rows [doff, doff + dloc) of the output-space part
// of a GKL u vector).  The basis — the O(k^2 n) re-orthogonalisation traffic that bounds SLQ — is cut S ways; the operator is applied
// REPLICATED on the all-gathered vector (the probe-batched mat-vec at B <= 4 is latency-bound, so sharding it would not shorten it).
// Per step: one all-gather of the new Krylov vector (+ one of its small output-space part for GKL) and one all-reduce per scalar /
// coefficient vector; all enqueued on the caller's stream between the kernels that produce and consume them.
struct Shard {
  lip_comm* comm = nullptr;
  const NcclApi* api = nullptr;
  int S = 1, rank = 0;
  int64_t Dsh = 0, off = 0, nloc = 0;    // D partition: shard width (multiple of 4), this rank's offset and length
  int64_t dsh = 0, doff = 0, dloc = 0;   // d partition (GKL kind)
};

int shard_init(Shard& sh, lip_comm* comm, int64_t D, int64_t d) {
  sh.comm = comm;
  sh.api = nccl_api();
  if (!sh.api) return LIP_ERR_UNSUPPORTED;
  sh.S = comm->world; sh.rank = comm->rank;
  sh.Dsh = pad4(ceil_div(D, sh.S));
  sh.off = sh.rank * sh.Dsh;
  sh.nloc = std::min<int64_t>(sh.Dsh, D - sh.off);
  LIP_REQUIRE((int64_t)(sh.S - 1) * sh.Dsh < D, "sharded Krylov: %d ranks are too many for vectors of length %lld", sh.S, (long long)D);
  if (d > 0) {
    sh.dsh = pad4(ceil_div(d, sh.S));
    sh.doff = std::min<int64_t>(sh.rank * sh.dsh, d);
    sh.dloc = std::max<int64_t>(0, std::min<int64_t>(sh.dsh, d - sh.doff));
  }
  return LIP_OK;
}

int shard_allreduce(const Shard* sh, float* buf, size_t count, cudaStream_t st) {
  if (!sh) return LIP_OK;
  LIP_CHECK_NCCL(sh->api, sh->api->AllReduce(buf, buf, count, ncclFloat, ncclSum, sh->comm->comm, st));
  count_launch();
  return LIP_OK;
}

// rank-local sums of squares -> global norms
int shard_norm(const Shard* sh, float* ssq, int64_t B, cudaStream_t st) {
  if (!sh) return LIP_OK;
  int rc = shard_allreduce(sh, ssq, (size_t)B, st);
  if (rc) return rc;
  sqrt_kernel<<<(unsigned)ceil_div(B, 128), 128, 0, st>>>(ssq, (int)B);
  LIP_LAUNCH_CHECK();
  return LIP_OK;
}

// loc [B][cnt] (this rank's zero-padded slice of every row) -> full [B][n] contiguous on every rank, n <= S * cnt
int shard_allgather(const Shard* sh, const float* loc, int64_t cnt, float* gath, float* full, int64_t n, int64_t B, cudaStream_t st) {
  const int64_t wide = (int64_t)sh->S * cnt;
  float* dst = (wide == n) ? full : gath;
  LIP_CHECK_NCCL(sh->api, sh->api->GroupStart());
  for (int64_t b = 0; b < B; ++b)
    LIP_CHECK_NCCL(sh->api, sh->api->AllGather(loc + b * cnt, dst + b * wide, (size_t)cnt, ncclFloat, sh->comm->comm, st));
  LIP_CHECK_NCCL(sh->api, sh->api->GroupEnd());
  count_launch();
  if (dst != full)
    LIP_CHECK_CUDA(cudaMemcpy2DAsync(full, sizeof(float) * n, gath, sizeof(float) * wide, sizeof(float) * n, (size_t)B,
                                     cudaMemcpyDeviceToDevice, st));
  return LIP_OK;
}

// ---- Lanczos (matfree decomp.tridiag_sym, reortho="full") -------------------------------------------------------------
size_t lanczos_ws_bytes(const Op& o, int64_t k, int64_t B) {
  const int64_t n = o.n_in, ld = pad4(n);
  return op_ws_bytes(o, B) + red_bytes(n, B, k) + 3 * rsz((size_t)B * (ld + 64), 4) + 2 * rsz((size_t)B * n, 4) + 8 * rsz((size_t)B + 4, 4) + 8192;
}

// sh == nullptr: the whole vectors live on this GPU.  Otherwise Q holds this rank's column slice [B, k, ldq >= nloc] and v0 is the
// FULL start vector (identical on every rank).
int lanczos_run(Op& o, const Shard* sh, const float* v0, int64_t ldv0, int64_t k, int64_t B, int passes, float* Q, int64_t ldq,
                float* diag, float* off, float* norm0, void* ws, size_t ws_bytes, cudaStream_t st) {
  const int64_t n = o.n_in;
  LIP_REQUIRE(B <= MAX_COLUMNS, "lanczos_run: at most 65535 probe columns per call, got %lld", (long long)B);
  LIP_REQUIRE(o.symmetric || o.n_in == o.n_out, "lanczos: the operator must be square");
  LIP_REQUIRE(o.op->kind != LIP_LINOP_GKL, "lanczos: the GKL operator is rectangular (use lip_gkl_bidiag)");
  LIP_REQUIRE(k >= 1 && k <= n, "num_matvecs=%lld exceeds the operator dimension %lld", (long long)k, (long long)n);
  LIP_REQUIRE(passes == 1 || passes == 2, "lanczos: passes must be 1 or 2");
  const int64_t nl = sh ? sh->nloc : n;            // columns this rank owns
  const int64_t c0 = sh ? sh->off : 0;             // ... starting at
  LIP_REQUIRE(Q && al16(Q) && ldq % 4 == 0 && ldq >= nl, "lanczos: basis must be 16-byte aligned with ldq %% 4 == 0 and ldq >= n");
  int rc = set_kernel_limits(k);
  if (rc) return rc;
  Bump bp(ws, ws_bytes);
  op_carve(o, bp, B);
  Red r;
  rc = red_carve(r, bp, nl, B, k, st);
  if (rc) return rc;
  r.sq_out = sh ? 1 : 0;
  const int64_t ld = pad4(nl);
  const int64_t lq = sh ? sh->Dsh : n;               // row stride of the contiguous / gather-source copy of q
  float* w = bp.take<float>((size_t)B * ld);         // working vector (this rank's slice), padded rows
  float* qloc = sh ? bp.take<float>((size_t)B * lq) : nullptr;             // this rank's slice of q, zero-padded to the shard width
  float* gath = sh ? bp.take<float>((size_t)B * sh->S * sh->Dsh) : nullptr;
  float* qc = bp.take<float>((size_t)B * n);         // contiguous FULL mat-vec input
  float* wc = bp.take<float>((size_t)B * n);         // contiguous FULL mat-vec output
  float* len = bp.take<float>((size_t)B);
  float* nrm0 = bp.take<float>((size_t)B);
  float* nrm1 = bp.take<float>((size_t)B);           // three-term form: norm after the two-row pass / after a gated extra pass
  float* len2 = bp.take<float>((size_t)B);
  int* gate = bp.take<int>((size_t)B + 4);
  if (!bp.ok) { set_error("lanczos: workspace too small (%zu bytes given)", ws_bytes); return LIP_ERR_WORKSPACE; }
  LIP_CHECK_CUDA(cudaMemsetAsync(len2, 0, sizeof(float) * (size_t)B, st));
  const int64_t qsb = k * ldq;
  if (ld != nl) LIP_CHECK_CUDA(cudaMemsetAsync(w, 0, sizeof(float) * (size_t)B * ld, st));
  if (sh) LIP_CHECK_CUDA(cudaMemsetAsync(qloc, 0, sizeof(float) * (size_t)B * lq, st));
  // q_0 = v0 / |v0|   (the basis rows must have zero padding: they are written column-exact into zeroed memory)
  LIP_CHECK_CUDA(cudaMemsetAsync(Q, 0, sizeof(float) * (size_t)B * qsb, st));
  float* qdst = sh ? qloc : qc;
  {
    AxpyNormArgs a{}; a.x1 = v0 + c0; a.ld1 = ldv0; a.n1 = nl; a.s1 = 1.f; a.out = w; a.ldo = ld; a.nrm = nrm0; a.n = nl;
    rc = launch_axpy_norm(a, B, r, st); if (rc) return rc;
    rc = shard_norm(sh, nrm0, B, st); if (rc) return rc;
    if (norm0) LIP_CHECK_CUDA(cudaMemcpyAsync(norm0, nrm0, sizeof(float) * B, cudaMemcpyDeviceToDevice, st));
    ScaleStoreArgs s{}; s.x = w; s.ldx = ld; s.scal = nrm0; s.o1 = Q; s.ld1 = ldq; s.o1sb = qsb; s.o2 = qdst; s.ld2 = lq; s.n2 = nl; s.n = nl;
    rc = launch_scale_store(s, B, st); if (rc) return rc;
  }
  const unsigned gB = (unsigned)ceil_div(B, 128);
  static const int three_term = getenv("LIP_LANCZOS_3TERM") ? atoi(getenv("LIP_LANCZOS_3TERM")) : 1;
  for (int64_t i = 0; i < k; ++i) {
    const int kk = (int)(i + 1);
    if (sh) { rc = shard_allgather(sh, qloc, sh->Dsh, gath, qc, n, B, st); if (rc) return rc; }
    rc = op_apply(o, qc, wc, B, 0, st); if (rc) return rc;
    const float* wl = wc + c0;                      // this rank's slice of A q_i (row stride n)
    if (passes == 2 && three_term) {
      // Two CGS passes, the first one restricted to the two rows that carry everything but rounding: A q_i - (q_i^T A q_i) q_i -
      // (q_{i-1}^T A q_i) q_{i-1} is the three-term recurrence, after which A q_i has only eps |A| left along the older vectors, and
      // the ONE full pass that follows removes that without cancellation.  Same tridiagonal (the recorded coefficients are rows i - 1
      // and i of the first pass in both forms), same orthogonality, half the basis traffic of two full passes.  Where the full pass
      // still removes more than half of the vector - a breakdown: A q_i lies in span(Q), e.g. the sampler's Lanczos that runs past the
      // numerical rank of the Gram - the column takes a second full pass (decided on the device, per column).
      const int64_t j0 = i > 0 ? i - 1 : 0;
      const int k2 = (int)(i - j0 + 1);
      rc = launch_project(Q + j0 * ldq, ldq, qsb, k2, wl, n, nl, B, r, st); if (rc) return rc;
      rc = shard_allreduce(sh, r.h, (size_t)B * r.kpad, st); if (rc) return rc;
      lanczos_record_kernel<<<gB, 128, 0, st>>>(diag, off, r.h - j0, r.kpad, len, (int)i, (int)k, (int)B);
      LIP_LAUNCH_CHECK();
      rc = launch_subtract(Q + j0 * ldq, ldq, qsb, k2, wl, n, nullptr, w, ld, nrm1, nl, B, r, st); if (rc) return rc;
      rc = shard_norm(sh, nrm1, B, st); if (rc) return rc;
      rc = launch_project(Q, ldq, qsb, kk, w, ld, nl, B, r, st); if (rc) return rc;
      rc = shard_allreduce(sh, r.h, (size_t)B * r.kpad, st); if (rc) return rc;
      rc = launch_subtract(Q, ldq, qsb, kk, w, ld, nullptr, w, ld, len, nl, B, r, st); if (rc) return rc;
      rc = shard_norm(sh, len, B, st); if (rc) return rc;
      ratio_gate_kernel<<<gB, 128, 0, st>>>(gate, len, nrm1, 0.5f, (int)B);
      LIP_LAUNCH_CHECK();
      rc = launch_project(Q, ldq, qsb, kk, w, ld, nl, B, r, st, gate); if (rc) return rc;
      rc = shard_allreduce(sh, r.h, (size_t)B * r.kpad, st); if (rc) return rc;
      rc = launch_subtract(Q, ldq, qsb, kk, w, ld, nullptr, w, ld, len2, nl, B, r, st, gate); if (rc) return rc;
      rc = shard_norm(sh, len2, B, st); if (rc) return rc;
      gate_merge_kernel<<<gB, 128, 0, st>>>(len, len2, gate, (int)B);
      LIP_LAUNCH_CHECK();
      if (i + 1 < k) {
        ScaleStoreArgs s{}; s.x = w; s.ldx = ld; s.scal = len; s.o1 = Q + (i + 1) * ldq; s.ld1 = ldq; s.o1sb = qsb; s.o2 = qdst; s.ld2 = lq;
        s.n2 = nl; s.n = nl;
        rc = launch_scale_store(s, B, st); if (rc) return rc;
      }
      continue;
    }
    // two CGS passes against Q[0..i]; the first-pass coefficients are the Arnoldi column H[:, i]
    rc = launch_project(Q, ldq, qsb, kk, wl, n, nl, B, r, st); if (rc) return rc;
    rc = shard_allreduce(sh, r.h, (size_t)B * r.kpad, st); if (rc) return rc;
    lanczos_record_kernel<<<gB, 128, 0, st>>>(diag, off, r.h, r.kpad, len, (int)i, (int)k, (int)B);
    LIP_LAUNCH_CHECK();
    rc = launch_subtract(Q, ldq, qsb, kk, wl, n, nullptr, w, ld, passes == 1 ? len : nullptr, nl, B, r, st); if (rc) return rc;
    if (passes == 2) {
      rc = launch_project(Q, ldq, qsb, kk, w, ld, nl, B, r, st); if (rc) return rc;
      rc = shard_allreduce(sh, r.h, (size_t)B * r.kpad, st); if (rc) return rc;
      rc = launch_subtract(Q, ldq, qsb, kk, w, ld, nullptr, w, ld, len, nl, B, r, st); if (rc) return rc;
    }
    rc = shard_norm(sh, len, B, st); if (rc) return rc;
    if (i + 1 < k) {
      ScaleStoreArgs s{}; s.x = w; s.ldx = ld; s.scal = len; s.o1 = Q + (i + 1) * ldq; s.ld1 = ldq; s.o1sb = qsb; s.o2 = qdst; s.ld2 = lq;
      s.n2 = nl; s.n = nl;
      rc = launch_scale_store(s, B, st); if (rc) return rc;
    }
  }
  return LIP_OK;
}

// ---- Golub-Kahan-Lanczos (matfree decomp.bidiag, full re-orthogonalisation of both bases) -------------------------------
size_t gkl_ws_bytes(const Op& o, int64_t k, int64_t B) {
  const int64_t nc = o.n_in, nr = o.n_out, ldu = pad4(nr), ldv = pad4(nc);
  return op_ws_bytes(o, B) + red_bytes(std::max(nc, nr), B, k) + rsz((size_t)B * ldu, 4) + rsz((size_t)B * ldv, 4) +
         4 * rsz((size_t)B * (nc + 64), 4) + 3 * rsz((size_t)B * (nr + 64), 4) + 6 * rsz((size_t)B, 4) + 8192 +
         rsz((size_t)B * pad4(k), 4) + 3 * rsz((size_t)B + 4, 4);      // reduced u basis: unit vectors, second-pass gates, dead flags
}

// sh == nullptr: whole vectors.  Otherwise (GKL kind only) Us / Vs hold this rank's slices: a u vector is stored as
// [its nloc parameter-space columns | its dloc output-space rows], a v vector as its nloc columns; v0 is the FULL start vector.
//
// reduced (LIP_LINOP_GKL only; the u basis is not returned): for A = [sqrt(alpha) I; W^T] the parameter-space block of every u vector
// lies in span(v_0 .. v_i) - u_i[:D] = V c_i - so u_i is carried as the SHORT vector [c_i (k coefficients) ; u_i[D:] (d entries)].
// With V orthonormal (it is re-orthogonalised every step) the Euclidean inner product of the short vectors IS the inner product of
// the long ones, A v_i = [sqrt(alpha) e_i ; W^T v_i], and since A^T u_i - alpha_i v_i is re-orthogonalised against V anyway,
//   v_{i+1} beta_{i+1} = (I - V V^T)(A^T u_i - alpha_i v_i) = (I - V V^T)(W u_i[D:] - alpha_i v_i)
// (the sqrt(alpha) V c_i term is annihilated by the projection), with beta_{i+1} read AFTER the projection.  Same recurrence in exact
// arithmetic, same rounding level in fp32, but the O(k^2 (D + d)) re-orthogonalisation traffic of the u basis - half of the HBM bytes
// of a GKL logdet, and half of its memory - becomes O(k^2 (k + d)): L2-resident.  In a sharded run the short u side is simply
// replicated on every rank (no collective).
int gkl_run(Op& o, const Shard* sh, const float* v0, int64_t ldv0, int64_t k, int64_t B, float* Us, int64_t ldu, float* Vs, int64_t ldv,
            float* alphas, float* betas, float* norm0, void* ws, size_t ws_bytes, cudaStream_t st, bool reduced = false) {
  const int64_t nc = o.n_in, nr = o.n_out;
  const bool gkl = o.op->kind == LIP_LINOP_GKL;
  LIP_REQUIRE(!reduced || gkl, "gkl_run: the reduced u basis needs the LIP_LINOP_GKL operator");
  const int64_t kp = pad4(k);
  LIP_REQUIRE(B <= MAX_COLUMNS, "gkl_run: at most 65535 probe columns per call, got %lld", (long long)B);
  LIP_REQUIRE(!sh || gkl, "sharded gkl: only the LIP_LINOP_GKL operator is supported");
  LIP_REQUIRE(k >= 1 && k <= std::min(nc, nr), "num_matvecs=%lld exceeds the operator dimensions (%lld, %lld)", (long long)k,
              (long long)nr, (long long)nc);
  const int64_t D = o.D, d = o.d;
  const int64_t ncl = sh ? sh->nloc : nc;                   // local length of a v vector
  const int64_t nrl = reduced ? kp + d : (sh ? sh->nloc + sh->dloc : nr);        // local length of a u vector
  const int64_t c0 = sh ? sh->off : 0, d0 = sh ? sh->doff : 0;
  LIP_REQUIRE(Us && Vs && al16(Us) && al16(Vs) && ldu % 4 == 0 && ldv % 4 == 0 && ldu >= nrl && ldv >= ncl,
              "gkl: bases must be 16-byte aligned with leading dimensions that are multiples of 4");
  int rc = set_kernel_limits(k);
  if (rc) return rc;
  Bump bp(ws, ws_bytes);
  op_carve(o, bp, B);
  Red r;
  rc = red_carve(r, bp, std::max(ncl, nrl), B, k, st);
  if (rc) return rc;
  r.sq_out = sh ? 1 : 0;
  Red ru = r;                                         // reduction scratch / collectives of the u side
  const Shard* shu = reduced ? nullptr : sh;
  if (reduced) ru.sq_out = 0;
  const int64_t lu = pad4(nrl), lv = pad4(ncl);
  float* ebuf = reduced ? bp.take<float>((size_t)B * kp) : nullptr;   // sqrt(alpha) e_i
  int* gate = reduced ? bp.take<int>((size_t)B + 4) : nullptr;        // [B] second-pass gates + [1] how many fired
  int* dead = reduced ? bp.take<int>((size_t)B + 4) : nullptr;        // [B] columns whose Krylov space is exhausted
  float* nv2 = reduced ? bp.take<float>((size_t)B) : nullptr;
  float* u = bp.take<float>((size_t)B * lu);          // working u (local slice), padded rows
  float* v = bp.take<float>((size_t)B * lv);          // working v (local slice), padded rows
  float* vc = bp.take<float>((size_t)B * nc);         // contiguous FULL A input
  float* wv = bp.take<float>((size_t)B * nc);         // contiguous FULL A^T output
  float* uc = bp.take<float>((size_t)B * nr);         // contiguous A^T input   (GKL kind: [B, D] part then [B, d] part, both FULL width)
  float* tu = bp.take<float>((size_t)B * nr);         // contiguous A output    (GKL kind: only the [B, d] output-space part)
  float* vloc = sh ? bp.take<float>((size_t)B * sh->Dsh) : nullptr;        // gather sources: zero-padded local slices
  float* udloc = sh ? bp.take<float>((size_t)B * std::max<int64_t>(sh->dsh, 4)) : nullptr;
  float* gath = sh ? bp.take<float>((size_t)B * sh->S * std::max(sh->Dsh, sh->dsh)) : nullptr;
  float* alpha = bp.take<float>((size_t)B);
  float* beta = bp.take<float>((size_t)B);
  float* nu = bp.take<float>((size_t)B);
  float* nv = bp.take<float>((size_t)B);
  float* nrm0 = bp.take<float>((size_t)B);
  if (!bp.ok) { set_error("gkl: workspace too small (%zu bytes given)", ws_bytes); return LIP_ERR_WORKSPACE; }
  const int64_t usb = k * ldu, vsb = k * ldv;
  if (lu != nrl) LIP_CHECK_CUDA(cudaMemsetAsync(u, 0, sizeof(float) * (size_t)B * lu, st));
  if (reduced) {
    LIP_CHECK_CUDA(cudaMemsetAsync(ebuf, 0, sizeof(float) * (size_t)B * kp, st));
    LIP_CHECK_CUDA(cudaMemsetAsync(gate, 0, sizeof(int) * ((size_t)B + 4), st));
    LIP_CHECK_CUDA(cudaMemsetAsync(dead, 0, sizeof(int) * ((size_t)B + 4), st));
    LIP_CHECK_CUDA(cudaMemsetAsync(nv2, 0, sizeof(float) * (size_t)B, st));
  }
  if (lv != ncl) LIP_CHECK_CUDA(cudaMemsetAsync(v, 0, sizeof(float) * (size_t)B * lv, st));
  LIP_CHECK_CUDA(cudaMemsetAsync(Us, 0, sizeof(float) * (size_t)B * usb, st));
  LIP_CHECK_CUDA(cudaMemsetAsync(Vs, 0, sizeof(float) * (size_t)B * vsb, st));
  LIP_CHECK_CUDA(cudaMemsetAsync(betas, 0, sizeof(float) * (size_t)B * k, st));
  if (sh) {
    LIP_CHECK_CUDA(cudaMemsetAsync(vloc, 0, sizeof(float) * (size_t)B * sh->Dsh, st));
    LIP_CHECK_CUDA(cudaMemsetAsync(udloc, 0, sizeof(float) * (size_t)B * std::max<int64_t>(sh->dsh, 4), st));
    LIP_CHECK_CUDA(cudaMemsetAsync(uc, 0, sizeof(float) * (size_t)B * nr, st));     // only this rank's columns of the [B, D] part are ever written
  }
  const float sa = gkl ? sqrtf(o.op->alpha) : 0.f;
  float* ucD = uc;                       // GKL kind: parameter-space part of u, [B, D]
  float* ucd = uc + (size_t)B * D;       //           output-space part, [B, d]
  // v_0 = v0 / |v0|
  {
    AxpyNormArgs a{}; a.x1 = v0 + c0; a.ld1 = ldv0; a.n1 = ncl; a.s1 = 1.f; a.out = v; a.ldo = lv; a.nrm = nrm0; a.n = ncl;
    rc = launch_axpy_norm(a, B, r, st); if (rc) return rc;
    rc = shard_norm(sh, nrm0, B, st); if (rc) return rc;
    if (norm0) LIP_CHECK_CUDA(cudaMemcpyAsync(norm0, nrm0, sizeof(float) * B, cudaMemcpyDeviceToDevice, st));
    ScaleStoreArgs s{}; s.x = v; s.ldx = lv; s.scal = nrm0; s.o1 = Vs; s.ld1 = ldv; s.o1sb = vsb; s.n2 = ncl; s.n = ncl;
    if (sh) { s.o2 = vloc; s.ld2 = sh->Dsh; } else { s.o2 = vc; s.ld2 = nc; }
    rc = launch_scale_store(s, B, st); if (rc) return rc;
  }
  const unsigned gB = (unsigned)ceil_div(B, 128);
  for (int64_t i = 0; i < k; ++i) {
    // ---- u = A v_i - beta_i u_{i-1};  alpha_i = |u|
    if (sh) { rc = shard_allgather(sh, vloc, sh->Dsh, gath, vc, nc, B, st); if (rc) return rc; }
    AxpyNormArgs a{};
    if (gkl) {
      rc = lip_wt_apply(o.op->model, vc, tu, B, o.op->scale, LIP_FACTOR_SQRT, o.mws, o.mws_bytes, st); if (rc) return rc;
      if (reduced) {
        unit_kernel<<<gB, 128, 0, st>>>(ebuf, kp, i, sa, (int)B);
        LIP_LAUNCH_CHECK();
        a.x1 = ebuf; a.ld1 = kp; a.n1 = kp; a.s1 = 1.f; a.x2 = tu; a.ld2 = d;
      } else {
        a.x1 = vc + c0; a.ld1 = D; a.n1 = ncl; a.s1 = sa; a.x2 = tu + d0; a.ld2 = d;
      }
    } else {
      rc = op_apply(o, vc, tu, B, 0, st); if (rc) return rc;
      a.x1 = tu; a.ld1 = nr; a.n1 = nr; a.s1 = 1.f;
    }
    if (i > 0) { a.y = Us + (i - 1) * ldu; a.ldy = ldu; a.ysb = usb; a.coef = beta; }
    a.out = u; a.ldo = lu; a.nrm = alpha; a.n = nrl;
    rc = launch_axpy_norm(a, B, ru, st); if (rc) return rc;
    rc = shard_norm(shu, alpha, B, st); if (rc) return rc;
    if (reduced) record_alive_kernel<<<gB, 128, 0, st>>>(alphas, k, i, alpha, dead, (int)B);
    else record_kernel<<<gB, 128, 0, st>>>(alphas, k, i, alpha, (int)B);
    LIP_LAUNCH_CHECK();
    // ---- u <- normalise, CGS against U[0..i-1], renormalise, store
    rc = launch_project(Us, ldu, usb, (int)i, u, lu, nrl, B, ru, st); if (rc) return rc;
    if (i > 0) { rc = shard_allreduce(shu, ru.h, (size_t)B * ru.kpad, st); if (rc) return rc; }
    rc = launch_subtract(Us, ldu, usb, (int)i, u, lu, alpha, u, lu, nu, nrl, B, ru, st); if (rc) return rc;
    rc = shard_norm(shu, nu, B, st); if (rc) return rc;
    {
      ScaleStoreArgs s{}; s.x = u; s.ldx = lu; s.scal = nu; s.o1 = Us + i * ldu; s.ld1 = ldu; s.o1sb = usb; s.n = nrl;
      if (reduced) { s.o2 = nullptr; s.n2 = kp; s.o3 = ucd; s.ld3 = d; }
      else if (gkl && sh) { s.o2 = ucD + c0; s.ld2 = D; s.n2 = ncl; s.o3 = udloc; s.ld3 = sh->dsh; }
      else if (gkl) { s.o2 = ucD; s.ld2 = D; s.n2 = D; s.o3 = ucd; s.ld3 = d; }
      else { s.o2 = uc; s.ld2 = nr; s.n2 = nr; }
      rc = launch_scale_store(s, B, st); if (rc) return rc;
    }
    if (i + 1 == k) break;              // the last v would not be used (matfree computes it; it does not enter B)
    // ---- w = A^T u_i - alpha_i v_i;  beta_{i+1} = |w|
    if (reduced) {
      rc = lip_w_apply(o.op->model, ucd, wv, B, o.op->scale, LIP_FACTOR_SQRT, nullptr, 0.f, o.mws, o.mws_bytes, st); if (rc) return rc;
    } else if (gkl) {
      if (sh) { rc = shard_allgather(sh, udloc, sh->dsh, gath, ucd, d, B, st); if (rc) return rc; }
      // sharded: ucD holds only this rank's columns (zeros elsewhere), so only this rank's columns of wv are meaningful - the ones used
      rc = lip_w_apply(o.op->model, ucd, wv, B, o.op->scale, LIP_FACTOR_SQRT, ucD, sa, o.mws, o.mws_bytes, st); if (rc) return rc;
    } else {
      rc = op_apply(o, uc, wv, B, 1, st); if (rc) return rc;
    }
    AxpyNormArgs c{}; c.x1 = wv + c0; c.ld1 = nc; c.n1 = ncl; c.s1 = 1.f; c.y = Vs + i * ldv; c.ldy = ldv; c.ysb = vsb; c.coef = alpha;
    c.out = v; c.ldo = lv; c.nrm = beta; c.n = ncl;
    rc = launch_axpy_norm(c, B, r, st); if (rc) return rc;
    rc = shard_norm(sh, beta, B, st); if (rc) return rc;
    if (!reduced) {
      record_kernel<<<gB, 128, 0, st>>>(betas, k, i + 1, beta, (int)B);
      LIP_LAUNCH_CHECK();
    }
    rc = launch_project(Vs, ldv, vsb, (int)(i + 1), v, lv, ncl, B, r, st); if (rc) return rc;
    rc = shard_allreduce(sh, r.h, (size_t)B * r.kpad, st); if (rc) return rc;
    rc = launch_subtract(Vs, ldv, vsb, (int)(i + 1), v, lv, beta, v, lv, nv, ncl, B, r, st); if (rc) return rc;
    rc = shard_norm(sh, nv, B, st); if (rc) return rc;
    if (reduced) {
      // W u_i[D:] - alpha_i v_i still carries the sqrt(alpha) V c_i component that the explicit recurrence cancels by addition; where
      // the projection removed most of the vector (small beta_{i+1}, or a breakdown) the remainder is re-orthogonalised once more -
      // per column, decided on the device: the second-pass kernels return at once for the other columns
      reorth_gate_kernel<<<gB, 128, 0, st>>>(gate, nv, 0.5f, (int)B, gate + B);
      LIP_LAUNCH_CHECK();
      rc = launch_project(Vs, ldv, vsb, (int)(i + 1), v, lv, ncl, B, r, st, gate); if (rc) return rc;
      rc = shard_allreduce(sh, r.h, (size_t)B * r.kpad, st); if (rc) return rc;
      rc = launch_subtract(Vs, ldv, vsb, (int)(i + 1), v, lv, nullptr, v, lv, nv2, ncl, B, r, st, gate); if (rc) return rc;
      rc = shard_norm(sh, nv2, B, st); if (rc) return rc;
      gate_merge_kernel<<<gB, 128, 0, st>>>(nv, nv2, gate, (int)B);
      LIP_LAUNCH_CHECK();
      // beta_{i+1} = |(I - V V^T)(W u_i[D:] - alpha_i v_i)| = (norm before the projection) x (norm of the projected unit vector)
      beta_fix_kernel<<<gB, 128, 0, st>>>(betas, k, i + 1, beta, nv, dead, (int)B);
      LIP_LAUNCH_CHECK();
    }
    {
      ScaleStoreArgs s{}; s.x = v; s.ldx = lv; s.scal = nv; s.o1 = Vs + (i + 1) * ldv; s.ld1 = ldv; s.o1sb = vsb; s.n2 = ncl; s.n = ncl;
      if (sh) { s.o2 = vloc; s.ld2 = sh->Dsh; } else { s.o2 = vc; s.ld2 = nc; }
      rc = launch_scale_store(s, B, st); if (rc) return rc;
    }
  }
  if (reduced && getenv("LIP_GKL_DEBUG")) {      // diagnostics only (synchronises): how many (step, column) pairs took the second pass
    int fired = 0;
    LIP_CHECK_CUDA(cudaStreamSynchronize(st));
    LIP_CHECK_CUDA(cudaMemcpy(&fired, gate + B, sizeof(int), cudaMemcpyDeviceToHost));
    fprintf(stderr, "[lip gkl reduced] k=%lld B=%lld: second re-orthogonalisation pass on %d of %lld (step, column) pairs\n", (long long)k,
            (long long)B, fired, (long long)((k - 1) * B));
  }
  return LIP_OK;
}

// ---- Hutch++ v2 --------------------------------------------------------------------------------------------------------------------
int launch_gram(const float* A, int64_t lda, int sa, const float* Bm, int64_t ldb, int sb, int64_t n, double* part, double* out,
                float* outf, cudaStream_t st) {
  const int tiles = (int)(ceil_div(sa, GT) * ceil_div(sb, GT));
  int nsplit = (int)std::min<int64_t>(64, std::max<int64_t>(1, (3 * 148) / tiles));
  int64_t cols = ceil_div(ceil_div(n, nsplit), GC) * GC;
  nsplit = (int)ceil_div(n, cols);
  dim3 grid((unsigned)tiles, (unsigned)nsplit);
  gram_partial_kernel<<<grid, GT * GT, 0, st>>>(A, lda, sa, Bm, ldb, sb, n, cols, part);
  LIP_LAUNCH_CHECK();
  const int64_t count = (int64_t)sa * sb;
  gram_finish_kernel<<<(unsigned)ceil_div(count, 256), 256, 0, st>>>(part, nsplit, count, out, outf);
  LIP_LAUNCH_CHECK();
  return LIP_OK;
}

int launch_rowdot(const float* A, int64_t lda, const float* Bm, int64_t ldb, int s, int64_t n, double scale, double* part, double* acc,
                  int accumulate, float* outf, cudaStream_t st) {
  const int np = 4 * 148;
  rowdot_partial_kernel<<<np, VT, 0, st>>>(A, lda, Bm, ldb, s, n, part);
  LIP_LAUNCH_CHECK();
  rowdot_finish_kernel<<<1, 32, 0, st>>>(part, np, scale, acc, accumulate, outf);
  LIP_LAUNCH_CHECK();
  return LIP_OK;
}

// Out[so, n] = T[so, si] * In[si, n]  (+ add_scale * Add)   through the fp32 SIMT GEMM
int small_times_tall(const float* T, int so, int si, const float* In, int64_t ldi, float* Out, int64_t ldo, int64_t n, float scale,
                     const float* Add, float add_scale, cudaStream_t st) {
  GemmProblem g;
  g.M = so; g.N = n; g.K = si; g.batch = 1;
  g.A1.ptr = T; g.A1.s0 = si; g.A1.s1 = 1;
  g.B1.ptr = In; g.B1.s0 = ldi; g.B1.s1 = 1;
  g.C = Out; g.c_sm = ldo;
  g.epi.scale = scale;
  if (Add) { g.epi.add = Add; g.epi.add_scale = add_scale; }
  return gemm_simt(g, st);
}

size_t hutchpp_ws_bytes(const Op& o, int64_t s1, int64_t s2) {
  const int64_t n = o.n_in, ld = pad4(n);
  if (s1 > n) s1 = n;
  const int64_t sm = std::max(s1, s2);
  return op_ws_bytes(o, sm) + 3 * rsz((size_t)s1 * ld, 4) + 2 * rsz((size_t)s2 * ld, 4) + rsz((size_t)64 * sm * sm, 8) +
         3 * rsz((size_t)sm * sm, 8) + rsz((size_t)4 * 148, 8) + 16384;
}

int hutchpp_run(Op& o, const float* probes, int64_t ldp, int64_t s1, int64_t s2, float* out, int32_t* info, void* ws, size_t ws_bytes,
                cudaStream_t st) {
  const int64_t n = o.n_in;
  LIP_REQUIRE(o.symmetric || o.n_in == o.n_out, "hutchpp: the operator must be square");
  LIP_REQUIRE(o.op->kind != LIP_LINOP_GKL, "hutchpp: the GKL operator is rectangular");
  LIP_REQUIRE(s1 >= 1 && s2 >= 1, "hutchpp: need s1 >= 1 and s2 >= 1 (s1=%lld, s2=%lld)", (long long)s1, (long long)s2);
  const bool full = s1 >= n;     // jnp.linalg.qr(reduced) of an [n, s1 >= n] block spans R^n: Q = I is that basis and the estimate is
  if (full) s1 = n;              // tr(X) exactly (stochtrace.py:118-135, tests/test_stochtrace.py:90-97), G_perp = 0
  LIP_REQUIRE(full || s1 <= 1024, "hutchpp: s1 = %lld exceeds 1024 (one-CTA Cholesky of the s1 x s1 Gram)", (long long)s1);
  const int64_t sm = std::max(s1, s2), ld = pad4(n);
  Bump bp(ws, ws_bytes);
  op_carve(o, bp, sm);
  float* Y = bp.take<float>((size_t)s1 * ld);        // X S^T, then scratch of the QR ping-pong
  float* Q = bp.take<float>((size_t)s1 * ld);
  float* XQ = bp.take<float>((size_t)s1 * ld);
  float* Gp = bp.take<float>((size_t)s2 * ld);
  float* XG = bp.take<float>((size_t)s2 * ld);
  double* gpart = bp.take<double>((size_t)64 * sm * sm);
  double* C = bp.take<double>((size_t)sm * sm);
  double* Linv = bp.take<double>((size_t)sm * sm);
  float* Tm = (float*)bp.take<double>((size_t)sm * sm);
  double* dpart = bp.take<double>((size_t)4 * 148);
  double* acc = bp.take<double>(2);
  int* flag = bp.take<int>(4);
  if (!bp.ok) { set_error("hutchpp: workspace too small (%zu bytes given, %zu needed)", ws_bytes, hutchpp_ws_bytes(o, s1, s2)); return LIP_ERR_WORKSPACE; }
  LIP_CHECK_CUDA(cudaMemsetAsync(flag, 0, sizeof(int) * 4, st));
  if (full) {
    eye_kernel<<<(unsigned)ceil_div(n * n, 256), 256, 0, st>>>(Q, n);
    LIP_LAUNCH_CHECK();
    int rcf = op_apply(o, Q, XQ, n, 0, st); if (rcf) return rcf;
    rcf = launch_rowdot(XQ, n, Q, n, (int)n, n, 1.0, dpart, acc, 0, out, st); if (rcf) return rcf;
    if (info) LIP_CHECK_CUDA(cudaMemcpyAsync(info, flag, sizeof(int), cudaMemcpyDeviceToDevice, st));
    return LIP_OK;
  }
  const float* S = probes;
  const float* G = probes + s1 * ldp;
  int rc;
  // rows must be contiguous [s, n] for the operator; probes with ldp != n are compacted into Q first
  const float* Sin = S;
  if (ldp != n) {
    LIP_CHECK_CUDA(cudaMemcpy2DAsync(Q, sizeof(float) * n, S, sizeof(float) * ldp, sizeof(float) * n, (size_t)s1, cudaMemcpyDeviceToDevice, st));
    Sin = Q;
  }
  // Y = X S^T  (stochtrace.py:121-122)
  float* Yc = XQ;                                     // contiguous [s1, n] result, re-laid with ld below when n % 4 != 0
  rc = op_apply(o, Sin, Yc, s1, 0, st); if (rc) return rc;
  // Q = qr(Y).Q (stochtrace.py:123) by shifted CholeskyQR, three passes: Gram in float64, Cholesky + triangular inverse in one CTA,
  // Q <- L^{-1} Y with the SIMT GEMM.  Any orthonormal basis of span(Y) gives the same estimate (tr(Q^T X Q) and the deflation
  // depend on the span only).
  const float* src = Yc; int64_t lds = n;
  float* bufs[2] = {Y, Q};
  for (int pass = 0; pass < 3; ++pass) {
    float* dst = bufs[pass & 1];
    rc = launch_gram(src, lds, (int)s1, src, lds, (int)s1, n, gpart, C, nullptr, st); if (rc) return rc;
    chol_inv_kernel<<<1, 1024, 0, st>>>(C, (int)s1, Tm, Linv, flag);
    LIP_LAUNCH_CHECK();
    rc = small_times_tall(Tm, (int)s1, (int)s1, src, lds, dst, n, n, 1.f, nullptr, 0.f, st); if (rc) return rc;
    src = dst; lds = n;
  }
  const float* Qf = src;                              // [s1, n] contiguous orthonormal rows (= bufs[0] = Y's storage)
  // tr(Q^T X Q)  (stochtrace.py:126-127)
  rc = op_apply(o, Qf, XQ, s1, 0, st); if (rc) return rc;
  rc = launch_rowdot(XQ, n, Qf, n, (int)s1, n, 1.0, dpart, acc, 0, nullptr, st); if (rc) return rc;
  // G_perp = G - (G Q) Q^T  (stochtrace.py:130-131): P = G Q^T-coefficients [s2, s1] in float64, then one GEMM with the -P Q + G epilogue
  const float* Gin = G;
  if (ldp != n) {
    LIP_CHECK_CUDA(cudaMemcpy2DAsync(XG, sizeof(float) * n, G, sizeof(float) * ldp, sizeof(float) * n, (size_t)s2, cudaMemcpyDeviceToDevice, st));
    Gin = XG;
  }
  rc = launch_gram(Gin, n, (int)s2, Qf, n, (int)s1, n, gpart, nullptr, Tm, st); if (rc) return rc;
  rc = small_times_tall(Tm, (int)s2, (int)s1, Qf, n, Gp, n, n, -1.f, Gin, 1.f, st); if (rc) return rc;
  // tr(G_perp X G_perp^T) / s2  (stochtrace.py:132-134)
  rc = op_apply(o, Gp, XG, s2, 0, st); if (rc) return rc;
  rc = launch_rowdot(Gp, n, XG, n, (int)s2, n, 1.0 / (double)s2, dpart, acc, 1, out, st); if (rc) return rc;
  if (info) LIP_CHECK_CUDA(cudaMemcpyAsync(info, flag, sizeof(int), cudaMemcpyDeviceToDevice, st));
  return LIP_OK;
}

}  // namespace

// =====================================================================================================================
extern "C" {

static size_t slq_ws_bytes(const Op& o, int form, int64_t k, int64_t B, int world);

size_t lip_krylov_workspace_bytes(const lip_linop* op, int32_t routine, int64_t k, int64_t B) {
  Op o;
  if (op_init(o, op) != LIP_OK || k <= 0 || B <= 0) return 0;
  const int64_t nc = o.n_in, nr = o.n_out;
  const size_t tri = lip_tridiag_scratch_bytes(k, B, 1) + 4 * rsz((size_t)B * k, 4);
  switch (routine) {
    case LIP_KRYLOV_LANCZOS: return lanczos_ws_bytes(o, k, B);
    case LIP_KRYLOV_GKL: return gkl_ws_bytes(o, k, B);
    case LIP_KRYLOV_SLQ_LANCZOS: return slq_ws_bytes(o, LIP_SLQ_LANCZOS, k, B, 1);
    case LIP_KRYLOV_SLQ_GKL: return slq_ws_bytes(o, LIP_SLQ_GKL, k, B, 1);
    case LIP_KRYLOV_FUNM: return lanczos_ws_bytes(o, k, B) + rsz((size_t)B * k * pad4(nc), 4) + tri + rsz((size_t)B * pad4(nc), 4) + 4096;
    case LIP_KRYLOV_HUTCHPP: return hutchpp_ws_bytes(o, k, B);
    case LIP_KRYLOV_APPLY: return op_ws_bytes(o, B) + rsz((size_t)B * (nr > nc ? nr : nc), 4) + 4096;
    case LIP_KRYLOV_CG:
      return op_ws_bytes(o, B) + 3 * rsz((size_t)B * nc, 4) + 4 * rsz((size_t)B, 4) + align_up(lip_dot_scratch_bytes(nc, B), 256) + 4096;
    default: return 0;
  }
}

int lip_linop_apply(const lip_linop* op, const float* in, float* out, int64_t B, int32_t transpose, void* workspace,
                    size_t workspace_bytes, lip_stream_t stream) {
  Op o;
  int rc = op_init(o, op);
  if (rc) return rc;
  LIP_REQUIRE(in && out && in != out && workspace && B > 0, "lip_linop_apply: bad argument");
  cudaStream_t st = (cudaStream_t)stream;
  Bump bp(workspace, workspace_bytes);
  op_carve(o, bp, B);
  float* tmp = (op->kind == LIP_LINOP_GKL) ? bp.take<float>((size_t)B * (o.D + o.d)) : nullptr;
  if (!bp.ok) {
    set_error("lip_linop_apply: workspace too small (%zu bytes given, %zu needed)", workspace_bytes,
              lip_krylov_workspace_bytes(op, LIP_KRYLOV_APPLY, 1, B));
    return LIP_ERR_WORKSPACE;
  }
  if (op->kind != LIP_LINOP_GKL) return op_apply(o, in, out, B, transpose, st);
  const float sa = sqrtf(op->alpha);
  const int64_t D = o.D, d = o.d;
  Red none;
  if (!transpose) {          // out = [sqrt(alpha) v ; scale W^T v]
    rc = lip_wt_apply(op->model, in, tmp, B, op->scale, LIP_FACTOR_SQRT, o.mws, o.mws_bytes, st);
    if (rc) return rc;
    AxpyNormArgs a{}; a.x1 = in; a.ld1 = D; a.n1 = D; a.s1 = sa; a.x2 = tmp; a.ld2 = d; a.out = out; a.ldo = D + d; a.n = D + d;
    return launch_axpy_norm(a, B, none, st);
  }
  float* uD = tmp;           // out = sqrt(alpha) u[:D] + scale W u[D:]
  float* ud = tmp + (size_t)B * D;
  ScaleStoreArgs s{}; s.x = in; s.ldx = D + d; s.o2 = uD; s.ld2 = D; s.n2 = D; s.o3 = ud; s.ld3 = d; s.n = D + d;
  rc = launch_scale_store(s, B, st);
  if (rc) return rc;
  return lip_w_apply(op->model, ud, out, B, op->scale, LIP_FACTOR_SQRT, uD, sa, o.mws, o.mws_bytes, st);
}

int lip_lanczos_tridiag(const lip_linop* op, const float* v0, int64_t ldv0, int64_t k, int64_t B, int32_t passes, float* Q,
                        int64_t ldq, float* diag, float* off, float* norm0, void* workspace, size_t workspace_bytes,
                        lip_stream_t stream) {
  Op o;
  int rc = op_init(o, op);
  if (rc) return rc;
  LIP_REQUIRE(v0 && diag && (off || k == 1) && workspace && B > 0 && ldv0 >= o.n_in, "lip_lanczos_tridiag: bad argument");
  return lanczos_run(o, nullptr, v0, ldv0, k, B, passes, Q, ldq, diag, off, norm0, workspace, workspace_bytes, (cudaStream_t)stream);
}

int lip_gkl_bidiag(const lip_linop* op, const float* v0, int64_t ldv0, int64_t k, int64_t B, float* Us, int64_t ldu, float* Vs,
                   int64_t ldv, float* alphas, float* betas, float* norm0, void* workspace, size_t workspace_bytes,
                   lip_stream_t stream) {
  Op o;
  int rc = op_init(o, op);
  if (rc) return rc;
  LIP_REQUIRE(v0 && alphas && betas && workspace && B > 0 && ldv0 >= o.n_in, "lip_gkl_bidiag: bad argument");
  return gkl_run(o, nullptr, v0, ldv0, k, B, Us, ldu, Vs, ldv, alphas, betas, norm0, workspace, workspace_bytes, (cudaStream_t)stream);
}

// the GKL logdet of the structured operator carries its u basis in reduced form (gkl_run); LIP_GKL_REDUCED=0 keeps the explicit basis
static bool slq_reduced(const Op& o, int form) {
  static const int on = [] {
    const char* e = getenv("LIP_GKL_REDUCED");
    return (e && e[0] == '0') ? 0 : 1;
  }();
  return on && form == LIP_SLQ_GKL && o.op->kind == LIP_LINOP_GKL;
}

static size_t slq_ws_bytes(const Op& o, int form, int64_t k, int64_t B, int world) {
  const int64_t nc = o.n_in, nr = o.n_out;
  const int64_t ncl = pad4(ceil_div(nc, world)) + 4, nrl = ncl + pad4(ceil_div(std::max<int64_t>(o.d, 0), world)) + 4;
  const size_t tri = lip_tridiag_scratch_bytes(k, B, 0) + 4 * rsz((size_t)B * k, 4) + 2 * rsz((size_t)B, 4);
  size_t bases = rsz((size_t)B * k * pad4(world > 1 ? ncl : nc), 4);
  if (slq_reduced(o, form)) bases += rsz((size_t)B * k * pad4(pad4(k) + o.d), 4);
  else if (form == LIP_SLQ_GKL) bases += rsz((size_t)B * k * pad4(world > 1 ? nrl : nr), 4);
  return (form == LIP_SLQ_GKL ? gkl_ws_bytes(o, k, B) : lanczos_ws_bytes(o, k, B)) + bases + tri + 8192;
}

size_t lip_slq_workspace_bytes(const lip_linop* op, int32_t form, int64_t k, int64_t B, int32_t world) {
  Op o;
  if (op_init(o, op) != LIP_OK || k <= 0 || B <= 0 || world < 1) return 0;
  return slq_ws_bytes(o, form, k, B, world);
}

int lip_slq_quadrature_sharded(const lip_linop* op, lip_comm* comm, const float* probes, int64_t ldp, int64_t k, int64_t B, int32_t form,
                               int32_t fn, float clip_min, float* quad_out, void* workspace, size_t workspace_bytes,
                               lip_stream_t stream) {
  Op o;
  int rc = op_init(o, op);
  if (rc) return rc;
  LIP_REQUIRE(probes && quad_out && workspace && B > 0 && k > 0, "lip_slq_quadrature: bad argument");
  LIP_REQUIRE(form == LIP_SLQ_LANCZOS || form == LIP_SLQ_GKL, "lip_slq_quadrature: unknown form %d", form);
  cudaStream_t st = (cudaStream_t)stream;
  Shard shard;
  const Shard* sh = nullptr;
  if (comm && comm->world > 1) {
    LIP_REQUIRE(op->kind == LIP_LINOP_GGN || op->kind == LIP_LINOP_GKL, "sharded SLQ: the operator must be a model kind (GGN / GKL)");
    rc = shard_init(shard, comm, o.D, form == LIP_SLQ_GKL ? o.d : 0);
    if (rc) return rc;
    sh = &shard;
  }
  const int64_t nc = o.n_in, nr = o.n_out;
  const bool reduced = slq_reduced(o, form);
  const int64_t ldv = pad4(sh ? sh->nloc : nc), ldu = reduced ? pad4(pad4(k) + o.d) : pad4(sh ? sh->nloc + sh->dloc : nr);
  Bump bp(workspace, workspace_bytes);
  float* td = bp.take<float>((size_t)B * k);
  float* to = bp.take<float>((size_t)B * k);
  float* al = bp.take<float>((size_t)B * k);
  float* be = bp.take<float>((size_t)B * k);
  float* nrm = bp.take<float>((size_t)B);
  float* q = bp.take<float>((size_t)B);
  void* tsc = bp.take<char>(lip_tridiag_scratch_bytes(k, B, 0));
  float* Vs = bp.take<float>((size_t)B * k * ldv);
  float* Us = form == LIP_SLQ_GKL ? bp.take<float>((size_t)B * k * ldu) : nullptr;
  if (!bp.ok) {
    set_error("lip_slq_quadrature: workspace too small (%zu bytes given, %zu needed)", workspace_bytes,
              slq_ws_bytes(o, form, k, B, sh ? sh->S : 1));
    return LIP_ERR_WORKSPACE;
  }
  const size_t rest = (size_t)(bp.end - bp.p);
  if (form == LIP_SLQ_LANCZOS) {
    rc = lanczos_run(o, sh, probes, ldp, k, B, 2, Vs, ldv, td, to, nrm, bp.p, rest, st);
  } else {
    rc = gkl_run(o, sh, probes, ldp, k, B, Us, ldu, Vs, ldv, al, be, nrm, bp.p, rest, st, reduced);
    if (!rc && getenv("LIP_GKL_DEBUG")) {        // diagnostics only (synchronises): the bidiagonal of the first column
      std::vector<float> ha((size_t)k), hb((size_t)k);
      LIP_CHECK_CUDA(cudaStreamSynchronize(st));
      LIP_CHECK_CUDA(cudaMemcpy(ha.data(), al, sizeof(float) * k, cudaMemcpyDeviceToHost));
      LIP_CHECK_CUDA(cudaMemcpy(hb.data(), be, sizeof(float) * k, cudaMemcpyDeviceToHost));
      int64_t bad = -1;
      for (int64_t i = 0; i < k && bad < 0; ++i) if (!(ha[i] == ha[i]) || !(hb[i] == hb[i]) || ha[i] > 1e30f || hb[i] > 1e30f) bad = i;
      fprintf(stderr, "[lip gkl] reduced=%d k=%lld first non-finite bidiagonal entry: %lld\n", (int)reduced, (long long)k, (long long)bad);
      const bool all = getenv("LIP_GKL_DEBUG")[0] == '2';
      const int64_t lo = all ? 0 : (bad >= 0 ? std::max<int64_t>(0, bad - 4) : 0);
      const int64_t hi = all ? (bad >= 0 ? bad + 1 : k) : (bad >= 0 ? std::min<int64_t>(k, bad + 2) : std::min<int64_t>(k, 8));
      for (int64_t i = lo; i < hi; ++i) fprintf(stderr, "   i=%lld alpha=%.6e beta=%.6e\n", (long long)i, ha[i], hb[i]);
      if (bad < 0) for (int64_t i = std::max<int64_t>(8, k - 6); i < k; ++i) fprintf(stderr, "   i=%lld alpha=%.6e beta=%.6e\n", (long long)i, ha[i], hb[i]);
    }
    if (!rc) rc = lip_bidiag_to_tridiag(al, be, td, to, k, B, st);
  }
  if (rc) return rc;
  // GKL form: T = B^T B is positive semi-definite and matfree evaluates f on the squared SINGULAR VALUES of B (>= 0 by construction).
  // The eigenvalue route can return -1e-13 |T| for a zero singular value of the decoupled post-breakdown block (weight ~1e-13, but
  // log of it is NaN), so the spectrum is floored at a denormal-scale positive number, which is what the SVD route sees.
  const float clip = (form == LIP_SLQ_GKL && clip_min < 0.f) ? 1e-30f : clip_min;
  rc = lip_tridiag_funm(td, to, k, B, fn, clip, q, nullptr, nullptr, tsc, st);
  if (rc) return rc;
  quad_scale_kernel<<<(unsigned)ceil_div(B, 128), 128, 0, st>>>(quad_out, q, nrm, (int)B);
  LIP_LAUNCH_CHECK();
  return LIP_OK;
}

int lip_slq_quadrature(const lip_linop* op, const float* probes, int64_t ldp, int64_t k, int64_t B, int32_t form, int32_t fn,
                       float clip_min, float* quad_out, void* workspace, size_t workspace_bytes, lip_stream_t stream) {
  return lip_slq_quadrature_sharded(op, nullptr, probes, ldp, k, B, form, fn, clip_min, quad_out, workspace, workspace_bytes, stream);
}

int lip_funm_lanczos(const lip_linop* op, const float* v, int64_t ldv_in, int64_t k, int64_t B, int32_t fn, float clip_min,
                     const float* fn_params, float* out, int64_t ldo, void* workspace, size_t workspace_bytes, lip_stream_t stream) {
  Op o;
  int rc = op_init(o, op);
  if (rc) return rc;
  LIP_REQUIRE(v && out && workspace && B > 0 && k > 0 && ldo >= o.n_in, "lip_funm_lanczos: bad argument");
  cudaStream_t st = (cudaStream_t)stream;
  const int64_t n = o.n_in, ldq = pad4(n);
  Bump bp(workspace, workspace_bytes);
  float* td = bp.take<float>((size_t)B * k);
  float* to = bp.take<float>((size_t)B * k);
  float* fe1 = bp.take<float>((size_t)B * k);
  float* nrm = bp.take<float>((size_t)B);
  float* tmp = bp.take<float>((size_t)B * ldq);
  void* tsc = bp.take<char>(lip_tridiag_scratch_bytes(k, B, 1));
  float* Q = bp.take<float>((size_t)B * k * ldq);
  if (!bp.ok) {
    set_error("lip_funm_lanczos: workspace too small (%zu bytes given, %zu needed)", workspace_bytes,
              lip_krylov_workspace_bytes(op, LIP_KRYLOV_FUNM, k, B));
    return LIP_ERR_WORKSPACE;
  }
  rc = lanczos_run(o, nullptr, v, ldv_in, k, B, 2, Q, ldq, td, to, nrm, bp.p, (size_t)(bp.end - bp.p), st);
  if (rc) return rc;
  rc = lip_tridiag_funm_p(td, to, k, B, fn, clip_min, fn_params, nullptr, fe1, nullptr, tsc, st);
  if (rc) return rc;
  rowscale_kernel<<<(unsigned)ceil_div(B * k, 256), 256, 0, st>>>(fe1, k, nrm, (int)k, (int)B);      // |v| f(T) e1
  LIP_LAUNCH_CHECK();
  rc = lip_basis_combine(Q, ldq, k, k, fe1, k, tmp, ldq, n, B, st);
  if (rc) return rc;
  LIP_CHECK_CUDA(cudaMemcpy2DAsync(out, sizeof(float) * ldo, tmp, sizeof(float) * ldq, sizeof(float) * n, (size_t)B,
                                   cudaMemcpyDeviceToDevice, st));
  return LIP_OK;
}

int lip_hutchpp_v2(const lip_linop* op, const float* probes, int64_t ldp, int64_t s1, int64_t s2, float* out, int32_t* info,
                   void* workspace, size_t workspace_bytes, lip_stream_t stream) {
  Op o;
  int rc = op_init(o, op);
  if (rc) return rc;
  LIP_REQUIRE(probes && out && workspace && ldp >= o.n_in, "lip_hutchpp_v2: bad argument");
  return hutchpp_run(o, probes, ldp, s1, s2, out, info, workspace, workspace_bytes, (cudaStream_t)stream);
}

int lip_cg_solve(const lip_linop* op, const float* b, float* x, int64_t B, float tol, float atol, int64_t maxiter,
                 int32_t check_every, int32_t* iters_out, void* workspace, size_t workspace_bytes, lip_stream_t stream) {
  Op o;
  int rc = op_init(o, op);
  if (rc) return rc;
  LIP_REQUIRE(b && x && workspace && B > 0, "lip_cg_solve: bad argument");
  LIP_REQUIRE(o.symmetric && o.op->kind != LIP_LINOP_GKL, "lip_cg_solve: the operator must be symmetric");
  cudaStream_t st = (cudaStream_t)stream;
  const int64_t n = o.n_in;
  if (maxiter < 0) maxiter = 10 * n;          // jax.scipy.sparse.linalg.cg default
  Bump bp(workspace, workspace_bytes);
  op_carve(o, bp, B);
  float* r = bp.take<float>((size_t)B * n);
  float* p = bp.take<float>((size_t)B * n);
  float* Ap = bp.take<float>((size_t)B * n);
  float* gamma = bp.take<float>((size_t)B);
  float* thresh = bp.take<float>((size_t)B);
  int* active = bp.take<int>((size_t)B);
  int* iters = bp.take<int>((size_t)B);
  int* any = bp.take<int>(4);
  void* dsc = bp.take<char>(lip_dot_scratch_bytes(n, B));
  if (!bp.ok) {
    set_error("lip_cg_solve: workspace too small (%zu bytes given, %zu needed)", workspace_bytes,
              lip_krylov_workspace_bytes(op, LIP_KRYLOV_CG, 1, B));
    return LIP_ERR_WORKSPACE;
  }
  rc = lip_cg_init(b, x, r, p, gamma, thresh, active, iters, tol, atol, n, B, dsc, st);
  if (rc) return rc;
  // early exit: every check_every iterations the OR of the active flags is copied to pinned host memory behind an event; the
  // host looks at the PREVIOUS check (already complete or nearly so) and never blocks on the current one.  Iterations enqueued
  // after convergence are no-ops for converged columns (their `active` flag is 0), so running ahead is harmless.
  int* host_flag = nullptr;
  cudaEvent_t ev[2] = {nullptr, nullptr};
  if (check_every > 0) {
    LIP_CHECK_CUDA(cudaHostAlloc((void**)&host_flag, 2 * sizeof(int), cudaHostAllocDefault));
    host_flag[0] = host_flag[1] = 1;
    LIP_CHECK_CUDA(cudaEventCreateWithFlags(&ev[0], cudaEventDisableTiming));
    LIP_CHECK_CUDA(cudaEventCreateWithFlags(&ev[1], cudaEventDisableTiming));
  }
  int pending = -1, slot = 0;
  int64_t it = 0;
  int status = LIP_OK;
  for (; it < maxiter; ++it) {
    if (check_every > 0 && it % check_every == 0) {
      if (pending >= 0) {
        cudaError_t e = cudaEventSynchronize(ev[pending]);
        if (e != cudaSuccess) { set_error("lip_cg_solve: %s", cudaGetErrorString(e)); status = LIP_ERR_CUDA; break; }
        if (host_flag[pending] == 0) break;
      }
      any_active_kernel<<<1, 256, 0, st>>>(active, (int)B, any);
      count_launch();
      cudaMemcpyAsync(&host_flag[slot], any, sizeof(int), cudaMemcpyDeviceToHost, st);
      cudaEventRecord(ev[slot], st);
      pending = slot;
      slot ^= 1;
    }
    status = op_apply(o, p, Ap, B, 0, st);
    if (status) break;
    status = lip_cg_step(x, r, p, Ap, gamma, thresh, active, iters, n, B, dsc, st);
    if (status) break;
  }
  if (iters_out && status == LIP_OK) cudaMemcpyAsync(iters_out, iters, sizeof(int) * B, cudaMemcpyDeviceToDevice, st);
  if (check_every > 0) {
    cudaStreamSynchronize(st);       // the pinned flag and the events are released below
    cudaEventDestroy(ev[0]); cudaEventDestroy(ev[1]);
    cudaFreeHost(host_flag);
  }
  return status;
}

}  // extern "C"
