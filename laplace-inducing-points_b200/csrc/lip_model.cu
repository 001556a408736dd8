// lip_model.cu — model handle, cached forward pass and the probe-batched GGN / W / W^T operators.
//
// Reference semantics: src/ggn.py:9-146 (compute_W_vps, compute_ggn_vp), src/lla.py:11-23,150-154.
// B200 design: the reference re-runs 3 forwards + 1 tangent + 1 backward per point per probe with batch-1
// GEMVs inside a sequential fori_loop over the M inducing points.  Here the forward pass runs ONCE at
// bind time; a product is then 2 GEMM families over all (probe, point) pairs at once:
//   JVP   layer l:  dH_l[b] = A_{l-1} dW_l[b] + dA_{l-1}[b] W_l + db_l[b];   dA_l = phi'_l * dH_l
//   VJP   layer l:  gW_l[b] = A_{l-1}^T D_l[b];  gb_l[b] = colsum D_l[b];  D_{l-1}[b] = (D_l[b] W_l^T) * phi'_{l-1}
// with the output-space Hessian / sqrt-Hessian applied in registers between the two sweeps.
#include <vector>
#include <stdlib.h>
#include <new>

#include "lip_common.cuh"
#include "lip_model.cuh"

using namespace lip;

namespace {

// ---- small fused kernels ----------------------------------------------------------------------------------
__global__ void softmax_rows_kernel(const float* __restrict__ f, float* __restrict__ P, float* __restrict__ S,
                                    int64_t M, int K) {
  int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i >= M) return;
  const float* fi = f + i * K;
  float mx = fi[0];
  for (int k = 1; k < K; ++k) mx = fmaxf(mx, fi[k]);
  float sum = 0.f;
  for (int k = 0; k < K; ++k) sum += expf(fi[k] - mx);
  float inv = 1.f / sum;
  for (int k = 0; k < K; ++k) {
    float p = expf(fi[k] - mx) * inv;
    P[i * K + k] = p;
    S[i * K + k] = sqrtf(p);
  }
}

// mode 0: H u  = p*u - p (p.u)            (ggn.py:125-129)
// mode 1: L^T u = s*u - (p.u) s            (ggn.py:29-39, 'sqrt_Hi_apply')
// mode 2: L u   = s*u - (s.u) p            (ggn.py:16-27, 'sqrt_Hi_apply_T')
// rows = B*M rows of K entries; P,S indexed by row % M.  out may alias in.
__global__ void factor_rows_kernel(const float* __restrict__ in, float* __restrict__ out,
                                   const float* __restrict__ P, const float* __restrict__ S, int64_t rows,
                                   int64_t M, int K, int mode, float scale) {
  int64_t r = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (r >= rows) return;
  const int64_t i = r % M;
  const float* u = in + r * K;
  const float* p = P + i * K;
  const float* s = S + i * K;
  float dot = 0.f;
  if (mode == 2) { for (int k = 0; k < K; ++k) dot += s[k] * u[k]; }
  else           { for (int k = 0; k < K; ++k) dot += p[k] * u[k]; }
  for (int k = 0; k < K; ++k) {
    float v;
    if (mode == 0) v = p[k] * u[k] - p[k] * dot;
    else if (mode == 1) v = s[k] * u[k] - dot * s[k];
    else v = s[k] * u[k] - dot * p[k];
    out[r * K + k] = scale * v;
  }
}

__global__ void scale_copy_kernel(const float* __restrict__ in, float* __restrict__ out, int64_t n, float scale) {
  int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i < n) out[i] = scale * in[i];
}

// gb[b][j] = scale * sum_m (Delta[b][m][j] (+ Delta_lo[b][m][j])) + add_scale * add[b][j];  Delta rows have stride ld
__global__ void bias_grad_kernel(const float* __restrict__ Delta, const float* __restrict__ Delta_lo, int64_t M, int n,
                                 int64_t ld, float* __restrict__ out, int64_t out_sz, float scale,
                                 const float* __restrict__ add, int64_t add_sz, float add_scale) {
  const int b = blockIdx.y;
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= n) return;
  const float* d = Delta + (int64_t)b * M * ld + j;
  float acc = 0.f;
  int64_t m = 0;
  for (; m + 4 <= M; m += 4) {
    float a0 = d[(m + 0) * ld], a1 = d[(m + 1) * ld], a2 = d[(m + 2) * ld], a3 = d[(m + 3) * ld];
    acc += (a0 + a1) + (a2 + a3);
  }
  for (; m < M; ++m) acc += d[m * ld];
  if (Delta_lo) {
    const float* e = Delta_lo + (int64_t)b * M * ld + j;
    float acc2 = 0.f;
    for (m = 0; m < M; ++m) acc2 += e[m * ld];
    acc += acc2;
  }
  float v = scale * acc;
  if (add) v += add_scale * add[(int64_t)b * add_sz + j];
  out[(int64_t)b * out_sz + j] = v;
}

// The same for SMALL batches (Krylov recurrences push 1 - 8 probes): the one-thread-per-column kernel above is a serial chain of M
// strided loads per thread on a grid of a few CTAs (measured 67 us for M = 512 at B = 1).  Here a 32 x 32 block gives each column
// 32 row lanes (rows m = ty, ty + 32, ...: 16 loads per thread at M = 512) and adds the lanes in a fixed order.
__global__ void __launch_bounds__(1024) bias_grad_rows_kernel(const float* __restrict__ Delta, const float* __restrict__ Delta_lo,
                                                              int64_t M, int n, int64_t ld, float* __restrict__ out, int64_t out_sz,
                                                              float scale, const float* __restrict__ add, int64_t add_sz,
                                                              float add_scale) {
  __shared__ float sm[32][33];
  const int b = blockIdx.y, tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const int j = blockIdx.x * 32 + tx;
  float acc = 0.f;
  if (j < n) {
    const float* d = Delta + (int64_t)b * M * ld + j;
    int64_t m = ty;
    for (; m + 96 < M; m += 128) {
      const float a0 = d[m * ld], a1 = d[(m + 32) * ld], a2 = d[(m + 64) * ld], a3 = d[(m + 96) * ld];
      acc += (a0 + a1) + (a2 + a3);
    }
    for (; m < M; m += 32) acc += d[m * ld];
    if (Delta_lo) {
      const float* e = Delta_lo + (int64_t)b * M * ld + j;
      float acc2 = 0.f;
      for (m = ty; m < M; m += 32) acc2 += e[m * ld];
      acc += acc2;
    }
  }
  sm[ty][tx] = acc;
  __syncthreads();
  if (ty == 0 && j < n) {
    float t = 0.f;
#pragma unroll
    for (int y = 0; y < 32; ++y) t += sm[y][tx];
    float v = scale * t;
    if (add) v += add_scale * add[(int64_t)b * add_sz + j];
    out[(int64_t)b * out_sz + j] = v;
  }
}

// picks the row-parallel variant when the column-per-thread grid would leave most SMs idle
static inline void launch_bias_grad_any(const float* Delta, const float* Delta_lo, int64_t rows, int n, int64_t ld, int64_t bc, float* out,
                                        int64_t out_sz, float scale, const float* add, int64_t add_sz, float add_scale,
                                        cudaStream_t st) {
  if (rows >= 64 && lip::ceil_div(n, 128) * bc < 148) {       // measured: at 256 CTAs the column-per-thread form is the faster one
    dim3 grid((unsigned)lip::ceil_div(n, 32), (unsigned)bc);
    bias_grad_rows_kernel<<<grid, 1024, 0, st>>>(Delta, Delta_lo, rows, n, ld, out, out_sz, scale, add, add_sz, add_scale);
  } else {
    dim3 grid((unsigned)lip::ceil_div(n, 128), (unsigned)bc);
    bias_grad_kernel<<<grid, 128, 0, st>>>(Delta, Delta_lo, rows, n, ld, out, out_sz, scale, add, add_sz, add_scale);
  }
}

// Tall-skinny contiguous case (conv stages: rows = points x pixels, n = channels, ld == n): one CTA per batch entry, the
// thread count is a multiple of n so every thread always sees the same column of the flat [rows * n] array.
__global__ void colsum_flat_kernel(const float* __restrict__ Delta, long long rows, int n, float* __restrict__ out,
                                   long long out_sz, float scale, const float* __restrict__ add, long long add_sz,
                                   float add_scale) {
  extern __shared__ float sm[];
  const long long b = blockIdx.x;
  const float* d = Delta + b * rows * n;
  const long long total = rows * n;
  float acc = 0.f;
  for (long long i = threadIdx.x; i < total; i += blockDim.x) acc += __ldg(d + i);
  sm[threadIdx.x] = acc;
  __syncthreads();
  if ((int)threadIdx.x < n) {
    float t = 0.f;
    for (int k = threadIdx.x; k < (int)blockDim.x; k += n) t += sm[k];
    float v = scale * t;
    if (add) v += add_scale * add[b * add_sz + threadIdx.x];
    out[b * out_sz + threadIdx.x] = v;
  }
}

// Delta back-propagation through a narrow head layer (K = out <= 16 logits):
//   C[r][n] = mask[r % M][n] * sum_k D[r][k] * W[n][k],   r = (probe, point),  W = kernel [in, K] row-major
// one thread per output element; the whole kernel matrix sits in shared memory.  Replaces a 128x128-tile GEMM whose K loop
// had a single, mostly padded k-tile (158 us -> HBM-bound).
__global__ void __launch_bounds__(256) head_dgrad_kernel(const float* __restrict__ D, const float* __restrict__ W,
                                                         const float* __restrict__ mask, float* __restrict__ C,
                                                         float* __restrict__ C_lo, long long rows, long long M, int in, int K,
                                                         int ld_out) {
  extern __shared__ float ws[];   // [in][K]
  for (int i = threadIdx.x; i < in * K; i += blockDim.x) ws[i] = W[i];
  __syncthreads();
  const long long total = rows * in;
  for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total; idx += (long long)gridDim.x * blockDim.x) {
    const long long r = idx / in;
    const int n = (int)(idx - r * in);
    const float* d = D + r * K;
    const float* w = ws + n * K;
    float acc = 0.f;
    for (int k = 0; k < K; ++k) acc = fmaf(__ldg(d + k), w[k], acc);
    if (mask) acc *= __ldg(mask + (r % M) * in + n);
    const long long co = r * ld_out + n;
    if (C_lo) {
      const float h = tf32_round(acc);
      C[co] = h;
      C_lo[co] = tf32_round(acc - h);
    } else {
      C[co] = acc;
    }
  }
}

// ---- fused head of lip_ggn_vp -------------------------------------------------------------------------------------------------------
// For a narrow last layer (K <= 16 outputs, e.g. the 128 -> 10 logit layer of the MNIST MLP) everything between the last wide JVP GEMM
// and the first wide VJP GEMM is one kernel, one CTA per probe (all resident at once at 256 probes), PC points per pass:
//   t_i   = A_i dW[b] + dA_i[b] W + db[b]                 (head JVP; ggn.py:139-141)
//   D_i   = H_i t_i  (softmax Hessian rows; regressor: t_i)  (ggn.py:125-131)
//   gW[b] += A_i^T D_i,   gb[b] += D_i                     (head weight / bias gradient, + alpha V in the epilogue; ggn.py:143)
//   Dn_i  = (D_i W^T) * phi'_i                             (delta of the layer below, stored as the TF32 (hi, lo) pair its GEMMs read)
// Replaces five launches (row-per-thread head JVP 85 us, factor rows 8 us, folded head weight gradient 168 us, bias gradient 24 us,
// head delta back-propagation 126 us per 256-probe call) whose intermediates ([B, M, K] logit tangents twice) went through HBM.
constexpr int HEAD_PC = 32;      // points per pass
constexpr int HEAD_KP = 17;      // padded row of the [in][K] weight tiles (odd: conflict-free column access)
struct HeadArgs {
  const float* A;                           // [M, in] cached input activations of the head
  const float* dA; long long dA_sz; int dA_ld;   // [B][M][dA_ld] masked tangent of the head's input (plain fp32)
  const float* V; const float* theta; long long D;   // D: probe stride of V
  long long ldo, lda;                       // probe strides of out / add
  long long woff, boff;                     // head kernel [in, K] / bias [K] offsets in the flat vector
  int in, K; long long M;
  const float* P; const float* S;           // softmax rows [M, K] (null: regressor, H = identity)
  const float* mask;                        // phi' of the layer below at the bound points, [M, in]
  float* out; float scale; const float* add; float add_scale;
  float* Dn_hi; float* Dn_lo; long long Dn_sz; int Dn_ld;
};

__global__ void __launch_bounds__(256) head_fused_kernel(HeadArgs a) {
  extern __shared__ float sm[];
  const int in = a.in, K = a.K, lda = in + 1;
  float* dWs = sm;                                   // [in][KP]
  float* Ws = dWs + in * HEAD_KP;                    // [in][KP]
  float* As = Ws + in * HEAD_KP;                     // [PC][in + 1]
  float* dAs = As + HEAD_PC * lda;                   // [PC][in + 1]
  float* part = dAs + HEAD_PC * lda;                 // [8][PC][K]
  float* Ds = part + 8 * HEAD_PC * K;                // [PC][KP]
  const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
  const long long b = blockIdx.x;
  const float* Vb = a.V + b * a.D;
  for (int idx = t; idx < in * K; idx += 256) {
    const int k = idx / K, n = idx - k * K;
    dWs[k * HEAD_KP + n] = Vb[a.woff + idx];
    Ws[k * HEAD_KP + n] = a.theta[a.woff + idx];
  }
  // weight-gradient accumulators: pair p = t + 256 r  ->  (k, n) = (p / K, p % K)
  float gw[16];
#pragma unroll
  for (int r = 0; r < 16; ++r) gw[r] = 0.f;
  const int npairs = in * K, R = (npairs + 255) >> 8;
  float gb = 0.f;                                    // threads t < K
  const int ks = in >> 3;                            // k slice per warp in the JVP step (in % 8 == 0)
  const float* dAb = a.dA + b * a.dA_sz;
  float* Dh = a.Dn_hi + b * a.Dn_sz;
  float* Dl = a.Dn_lo ? a.Dn_lo + b * a.Dn_sz : nullptr;
  for (long long i0 = 0; i0 < a.M; i0 += HEAD_PC) {
    const int valid = (int)((a.M - i0) < HEAD_PC ? (a.M - i0) : HEAD_PC);
    __syncthreads();
    for (int idx = t; idx < HEAD_PC * in; idx += 256) {
      const int i = idx / in, k = idx - i * in;
      const bool ok = i < valid;
      As[i * lda + k] = ok ? __ldg(a.A + (i0 + i) * in + k) : 0.f;
      dAs[i * lda + k] = ok ? dAb[(i0 + i) * a.dA_ld + k] : 0.f;
    }
    __syncthreads();
    {  // head JVP partials: lane = point, warp = k slice
      float acc[16];
#pragma unroll
      for (int n = 0; n < 16; ++n) acc[n] = 0.f;
      const float* ar = As + lane * lda;
      const float* dr = dAs + lane * lda;
      for (int k = warp * ks; k < (warp + 1) * ks; ++k) {
        const float av = ar[k], dv = dr[k];
        const float* w1 = dWs + k * HEAD_KP;
        const float* w2 = Ws + k * HEAD_KP;
#pragma unroll
        for (int n = 0; n < 16; ++n)
          if (n < K) acc[n] = fmaf(av, w1[n], fmaf(dv, w2[n], acc[n]));
      }
#pragma unroll
      for (int n = 0; n < 16; ++n)
        if (n < K) part[(warp * HEAD_PC + lane) * K + n] = acc[n];
    }
    __syncthreads();
    for (int idx = t; idx < HEAD_PC * K; idx += 256) {      // sum the 8 slices in a fixed order, add the bias tangent
      const int i = idx / K, n = idx - i * K;
      float v = Vb[a.boff + n];
#pragma unroll
      for (int g = 0; g < 8; ++g) v += part[(g * HEAD_PC + i) * K + n];
      Ds[i * HEAD_KP + n] = (i < valid) ? v : 0.f;
    }
    __syncthreads();
    if (a.P && t < valid) {                                  // D = H t = p * t - p (p . t)   (one thread per point)
      const float* p = a.P + (i0 + t) * K;
      float* u = Ds + t * HEAD_KP;
      float dot = 0.f;
      for (int n = 0; n < K; ++n) dot = fmaf(__ldg(p + n), u[n], dot);
      for (int n = 0; n < K; ++n) { const float pn = __ldg(p + n); u[n] = pn * u[n] - pn * dot; }
    }
    __syncthreads();
#pragma unroll
    for (int r = 0; r < 16; ++r) {                           // head weight gradient
      const int pidx = t + (r << 8);
      if (r < R && pidx < npairs) {
        const int k = pidx / K, n = pidx - k * K;
        float s = gw[r];
        for (int i = 0; i < HEAD_PC; ++i) s = fmaf(As[i * lda + k], Ds[i * HEAD_KP + n], s);
        gw[r] = s;
      }
    }
    if (t < K) {
      float s = gb;
      for (int i = 0; i < HEAD_PC; ++i) s += Ds[i * HEAD_KP + t];
      gb = s;
    }
    for (int idx = t; idx < HEAD_PC * in; idx += 256) {      // delta of the layer below
      const int i = idx / in, k = idx - i * in;
      if (i >= valid) continue;
      const float* d = Ds + i * HEAD_KP;
      const float* w = Ws + k * HEAD_KP;
      float v = 0.f;
#pragma unroll
      for (int n = 0; n < 16; ++n)
        if (n < K) v = fmaf(d[n], w[n], v);
      v *= __ldg(a.mask + (i0 + i) * in + k);
      const long long co = (i0 + i) * a.Dn_ld + k;
      if (Dl) {
        const float h = tf32_round(v);
        Dh[co] = h;
        Dl[co] = tf32_round(v - h);
      } else {
        Dh[co] = v;
      }
    }
  }
  float* ob = a.out + b * a.ldo;
  const float* ab = a.add ? a.add + b * a.lda : nullptr;
#pragma unroll
  for (int r = 0; r < 16; ++r) {
    const int pidx = t + (r << 8);
    if (r < R && pidx < npairs) {
      float v = a.scale * gw[r];
      if (ab) v += a.add_scale * ab[a.woff + pidx];
      ob[a.woff + pidx] = v;
    }
  }
  if (t < K) {
    float v = a.scale * gb;
    if (ab) v += a.add_scale * ab[a.boff + t];
    ob[a.boff + t] = v;
  }
}

static inline size_t head_smem_bytes(int in, int K) {
  return sizeof(float) * ((size_t)2 * in * HEAD_KP + (size_t)2 * HEAD_PC * (in + 1) + (size_t)8 * HEAD_PC * K + (size_t)HEAD_PC * HEAD_KP);
}

__global__ void onehot_rows_kernel(float* __restrict__ U, int64_t d, int64_t start, int64_t blk) {
  int64_t idx = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (idx >= blk * d) return;
  int64_t j = idx / d, c = idx % d;
  U[idx] = (c == start + j) ? 1.f : 0.f;
}

// final[r][c] = G'[max(r,c)][min(r,c)]  where row c of G' holds WT(W(e_c)) (ggn.py:227 keeps the upper
// triangle of the column-filled matrix == the lower triangle of the row-filled one).
__global__ void symmetrize_from_lower_kernel(float* __restrict__ G, int64_t d) {
  int64_t idx = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (idx >= d * d) return;
  int64_t r = idx / d, c = idx % d;
  if (r < c) G[r * d + c] = G[c * d + r];
}

}  // namespace

namespace lip {
int launch_bias_grad(const float* Delta, const float* Delta_lo, int64_t rows, int n, int64_t ld, int64_t B, float* out,
                     int64_t out_sz, float scale, const float* add, int64_t add_sz, float add_scale, cudaStream_t st) {
  if (!Delta_lo && ld == n && n <= 64 && rows >= 2048) {
    const int threads = n * (1024 / n);
    colsum_flat_kernel<<<(unsigned)B, threads, threads * sizeof(float), st>>>(Delta, rows, n, out, out_sz, scale, add, add_sz,
                                                                              add_scale);
    LIP_LAUNCH_CHECK();
    return LIP_OK;
  }
  for (int64_t b0 = 0; b0 < B; b0 += 65535) {
    const int64_t bc = B - b0 < 65535 ? B - b0 : 65535;
    launch_bias_grad_any(Delta + b0 * rows * ld, Delta_lo ? Delta_lo + b0 * rows * ld : nullptr, rows, n, ld, bc,
                         out + b0 * out_sz, out_sz, scale, add ? add + b0 * add_sz : nullptr, add_sz, add_scale, st);
    LIP_LAUNCH_CHECK();
  }
  return LIP_OK;
}
int launch_scale_copy(const float* in, float* out, int64_t n, float scale, cudaStream_t st) {
  scale_copy_kernel<<<(unsigned)ceil_div(n, 256), 256, 0, st>>>(in, out, n, scale);
  LIP_LAUNCH_CHECK();
  return LIP_OK;
}
int launch_softmax(const float* logits, float* P, float* S, int64_t M, int K, cudaStream_t st) {
  softmax_rows_kernel<<<(unsigned)ceil_div(M, 128), 128, 0, st>>>(logits, P, S, M, K);
  LIP_LAUNCH_CHECK();
  return LIP_OK;
}

int launch_factor(const float* in, float* out, const lip_model* m, int64_t B, int mode, float scale,
                  cudaStream_t st) {
  int64_t rows = B * m->M;
  if (m->model_type == LIP_REGRESSOR) {
    // scalar factor: handled by the caller through `scale`
    int64_t n = rows * m->K;
    if (in != out || scale != 1.f) {
      scale_copy_kernel<<<(unsigned)ceil_div(n, 256), 256, 0, st>>>(in, out, n, scale);
      LIP_LAUNCH_CHECK();
    }
    return LIP_OK;
  }
  factor_rows_kernel<<<(unsigned)ceil_div(rows, 128), 128, 0, st>>>(in, out, m->P, m->S, rows, m->M, m->K, mode,
                                                                    scale);
  LIP_LAUNCH_CHECK();
  return LIP_OK;
}

}  // namespace lip

namespace {
static inline int pad4(int x) { return (x + 3) / 4 * 4; }

struct Workspace {
  float* hi[2];     // intermediate ping-pong (plain fp32, or the tf32-hi part when the consumer is a tcgen05 GEMM)
  float* lo[2];     // tf32-lo parts (tensor path only)
  float* vs_hi;     // split tangent-weight block of the current layer, [B][in][ldw]
  float* vs_lo;
  float* colsum;    // per-32-row-block column sums of the delta written by a tcgen05 delta-backprop GEMM: [B][nslots][ldmax]
  float* pre;       // one- / two-probe sweeps: mask * (A_l dW_l) of every layer, computed on the side stream, [layer][min(B,2)][M][ldmax]
  size_t per_buf;   // floats
  // row strides (floats) of the caller's blocks: V (JVP input / bias source), out (VJP output), add (VJP epilogue operand).
  // D unless the caller hands over padded rows (lip_ggn_vp_ex), which is what lets TMA read the probe block in place.
  int64_t ldv = 0, ldo = 0, lda = 0;
  bool exact = false;   // LIP_PROBES_EXACT_TF32: V's entries are exactly TF32-representable (+-1 probes, one-hot blocks)
};

static inline int64_t colsum_slots(const lip_model* m) { return ceil_div(m->M, 128) * 4; }

size_t ws_bytes(const lip_model* m, int64_t B) {
  size_t per = align_up((size_t)B * (size_t)m->M * (size_t)(m->tc_on ? m->ldmax : m->maxw), 64);
  size_t total = 2 * per;
  if (m->tc_on) total += 2 * per + 2 * align_up((size_t)B * (size_t)m->sum_split, 64) +
                        align_up((size_t)B * (size_t)colsum_slots(m) * (size_t)m->ldmax, 64) +
                        align_up((size_t)(B < 2 ? B : 2) * m->L.size() * (size_t)m->M * (size_t)m->ldmax, 64);
  return total * sizeof(float) + 256;
}

int carve(const lip_model* m, int64_t B, void* ws, size_t bytes, Workspace* w) {
  size_t need = ws_bytes(m, B);
  if (bytes < need || ws == nullptr) {
    set_error("workspace too small: need %zu bytes, got %zu", need, bytes);
    return LIP_ERR_WORKSPACE;
  }
  float* base = (float*)align_up((uintptr_t)ws, 256);
  size_t per = align_up((size_t)B * (size_t)m->M * (size_t)(m->tc_on ? m->ldmax : m->maxw), 64);
  w->per_buf = per;
  w->ldv = w->ldo = w->lda = m->D;
  w->exact = false;
  w->hi[0] = base; w->hi[1] = base + per;
  w->lo[0] = w->lo[1] = w->vs_hi = w->vs_lo = w->colsum = w->pre = nullptr;
  if (m->tc_on) {
    w->lo[0] = base + 2 * per; w->lo[1] = base + 3 * per;
    size_t vs = align_up((size_t)B * (size_t)m->sum_split, 64);
    w->vs_hi = base + 4 * per; w->vs_lo = w->vs_hi + vs;
    w->colsum = w->vs_lo + vs;
    w->pre = w->colsum + align_up((size_t)B * (size_t)colsum_slots(m) * (size_t)m->ldmax, 64);
  }
  return LIP_OK;
}

// leading dimension of the intermediate produced by layer l (its output width), padded on the tensor path
static inline int ld_of(const lip_model* m, int width) { return m->tc_on ? pad4(width) : width; }

// Layer l's probe block V[b, woff_l ...] can be the tcgen05 GEMM's B operand where it lies: the caller vouches that its entries are
// exactly TF32-representable (no lo part: +-1 Rademacher probes, one-hot blocks) and every TMA requirement holds - 16-byte aligned
// base, row stride `out` and probe stride ldv multiples of 4 floats.  (For the reference's [B, D] blocks with D % 4 != 0 - the
// MNIST MLP has D = 1,494,154 - the probe stride breaks this, which is why callers that own their probe buffers pad the rows.)
static inline bool inplace_ok(const lip_model* m, int l, const float* V, const Workspace& w) {
  const DenseLayer& Ld = m->L[l];
  return w.exact && m->tc_on && m->tc_layer[l] && (((uintptr_t)(V + Ld.woff)) & 15) == 0 && Ld.out % 4 == 0 && w.ldv % 4 == 0;
}

// the fused head kernel applies: dense program with >= 2 layers whose last layer is a narrow SIMT layer
static inline bool head_fusable(const lip_model* m) {
  static const bool off = getenv("LIP_HEAD_FUSE") && atoi(getenv("LIP_HEAD_FUSE")) == 0;
  const int nL = (int)m->L.size();
  if (off || nL < 2 || m->is_cnn || m->is_resnet) return false;
  if (m->tc_on && m->tc_layer[nL - 1]) return false;
  const DenseLayer& Lh = m->L[nL - 1];
  if (Lh.out > 16 || Lh.in > 256 || Lh.in % 8 != 0 || Lh.in < 8) return false;
  return head_smem_bytes(Lh.in, Lh.out) <= 200 * 1024;
}

// ---- JVP sweep: V[B,D] -> dlogits written to `dst` ([B,M,K], contiguous).  Intermediates ping-pong in ws.
// keep_hi / keep_lo (optional, [nL - 1] pointers): layer l < nL - 1 writes its masked tangent dA_{l+1} there instead of into the
// ping-pong buffers, so a later reverse pass (lip_zgrad.cu) can read every layer's tangent.
// stop_layer (default: all): run layers [0, stop_layer) only; with stop_layer = nL - 1 the masked tangent of the head's input is left in
// w.hi[(nL - 2) & 1] (plain fp32 when the head is a SIMT layer) for the fused head kernel.
int jvp_sweep(lip_model* m, const float* V, int64_t B, const Workspace& w, float* dst, cudaStream_t st,
              float* const* keep_hi = nullptr, float* const* keep_lo = nullptr, int stop_layer = -1, const float* t0_hi = nullptr,
              const float* t0_lo = nullptr, int t0_ld = 0) {
  const int nL = (int)m->L.size();
  const int nRun = stop_layer < 0 ? nL : stop_layer;
  // (t0_hi, t0_lo, t0_ld): tangent of the program's INPUT, [B, M, t0_ld] (a dense tail behind conv stages); null for a program whose
  // input is data
  const float* prev_hi = t0_hi;
  const float* prev_lo = t0_lo;
  int prev_ld = t0_ld;
  // TF32 (hi, lo) splits of the probes' weight blocks: HBM-bound, independent of everything but V.  They run on the
  // model's side stream, layer by layer, while the compute-bound GEMMs of the earlier layers run on `st`; the first
  // tensor-core layer is additionally cut into probe chunks so that only its first chunk's split is exposed.
  constexpr int NCH = lip_model::SPLIT_CHUNKS;
  // Measured on B200 (profiles/r01_split_overlap.txt): no gain - the sustained step is power-capped, so hiding the
  // HBM-bound splits behind the tensor-bound GEMMs does not shorten it.  Off unless LIP_SPLIT_OVERLAP=1.
  static const bool overlap = getenv("LIP_SPLIT_OVERLAP") && atoi(getenv("LIP_SPLIT_OVERLAP")) != 0;
  int first_tc = -1;
  for (int l = 0; l < nL && m->tc_on; ++l) if (m->tc_layer[l]) { first_tc = l; break; }
  const int64_t chunk = ceil_div(B, NCH);
  if (m->tc_on) {
    cudaStream_t ss = overlap ? m->side : st;
    if (overlap) {
      LIP_CHECK_CUDA(cudaEventRecord(m->ev_fork, st));
      LIP_CHECK_CUDA(cudaStreamWaitEvent(ss, m->ev_fork, 0));
    }
    LIP_CHECK_CUDA(cudaMemsetAsync(m->lo_nz, 0, sizeof(int) * nL, ss));
    for (int l = 0; l < nL; ++l) {
      if (!m->tc_layer[l]) continue;
      if (inplace_ok(m, l, V, w)) continue;     // exactly-TF32 probes in TMA-aligned rows: the GEMM reads V itself, no split pass
      const DenseLayer& Ld = m->L[l];
      const int64_t ldw = m->W_ld[l], bsz = (int64_t)Ld.in * ldw;
      float* hi = w.vs_hi + B * m->split_off[l];
      float* lo = w.vs_lo + B * m->split_off[l];
      const int nch = (l == first_tc && overlap) ? NCH : 1;
      for (int c = 0; c < nch; ++c) {
        const int64_t b0 = nch == 1 ? 0 : c * chunk, b1 = nch == 1 ? B : (b0 + chunk < B ? b0 + chunk : B);
        if (b1 > b0) {
          int rc = tf32_split3(V + b0 * w.ldv + Ld.woff, w.ldv, Ld.out, hi + b0 * bsz, lo + b0 * bsz, bsz, ldw, b1 - b0, Ld.in,
                               Ld.out, ss, m->lo_nz + l);
          if (rc) return rc;
        }
        if (overlap) LIP_CHECK_CUDA(cudaEventRecord(m->ev_split[l * NCH + c], ss));
      }
    }
  }
  // One or two probes: the sweep is a chain of latency-bound GEMMs, and half of every link - the A_l dW_l term - does not depend on
  // the chain at all.  Those products (already masked: the mask distributes over the sum) run on the side stream, all layers at once,
  // while the chain does T_l W_l + bias, masks, and adds them in its epilogue.  Same sums, K halved on the critical path.
  static const int pre_env = getenv("LIP_JVP_FORK") ? atoi(getenv("LIP_JVP_FORK")) : 1;
  const bool pre_fork = pre_env && m->tc_on && m->side != nullptr && B <= 2 && !overlap && w.pre != nullptr &&
                        (int)m->ev_split.size() >= nL * lip_model::SPLIT_CHUNKS && lip_model::SPLIT_CHUNKS >= 4;
  const size_t pre_stride = (size_t)B * (size_t)m->M * (size_t)m->ldmax;
  auto has_pre = [&](int l) {      // layers whose A dW term is split off: tensor-core layers with a tangent input, not the last one
    return pre_fork && l < nRun && l < nL - 1 && m->tc_layer[l] && (l > 0 || t0_hi != nullptr);
  };
  if (pre_fork) {
    bool any = false;
    for (int l = 0; l < nRun; ++l) any = any || has_pre(l);
    if (any) {
      LIP_CHECK_CUDA(cudaEventRecord(m->ev_fork, st));              // the probe splits above are on `st`
      LIP_CHECK_CUDA(cudaStreamWaitEvent(m->side, m->ev_fork, 0));
      for (int l = 0; l < nRun; ++l) {
        if (!has_pre(l)) continue;
        const DenseLayer& Ld = m->L[l];
        const int64_t ldw = m->W_ld[l];
        const int out_ld = ld_of(m, Ld.out);
        TcGemmProblem p;
        p.M = m->M; p.N = Ld.out; p.K = Ld.in; p.batch = B;
        p.A1.hi = m->A_hi[l]; p.A1.lo = m->A_lo[l]; p.A1.ld = m->A_ld[l]; p.A1.sz = m->M * m->A_ld[l]; p.A1.major_k = 1;
        p.a_batched = 0;
        p.B1.hi = w.vs_hi + B * m->split_off[l]; p.B1.lo = w.vs_lo + B * m->split_off[l]; p.B1.ld = ldw;
        p.B1.sz = (int64_t)Ld.in * ldw; p.B1.major_k = 0;
        if (inplace_ok(m, l, V, w)) { p.B1.hi = V + Ld.woff; p.B1.lo = V + Ld.woff; p.B1.ld = Ld.out; p.B1.sz = w.ldv; }
        p.b_batched = 1;
        p.B1.lo_nz = m->lo_nz + l;
        p.C = w.pre + l * pre_stride; p.C_lo = nullptr; p.c_sz = m->M * (int64_t)out_ld; p.c_sm = out_ld;
        p.epi.mask = m->dphi[l]; p.epi.mask_sm = Ld.out;
        int rc = gemm_tc(p, m->side);
        if (rc) return rc;
        LIP_CHECK_CUDA(cudaEventRecord(m->ev_split[l * lip_model::SPLIT_CHUNKS + 2], m->side));
      }
    }
  }
  for (int l = 0; l < nRun; ++l) {
    const DenseLayer& Ld = m->L[l];
    const bool last = (l == nL - 1);
    const bool tc = m->tc_on && m->tc_layer[l];
    const bool next_tc = m->tc_on && !last && m->tc_layer[l + 1];   // consumer of this layer's output
    float* out_hi = last ? dst : (keep_hi ? keep_hi[l] : w.hi[l & 1]);
    float* out_lo = (!last && next_tc) ? (keep_lo ? keep_lo[l] : w.lo[l & 1]) : nullptr;
    const int out_ld = last ? Ld.out : ld_of(m, Ld.out);
    if (tc) {
      const int64_t ldw = m->W_ld[l];
      const float* vs_hi = w.vs_hi + B * m->split_off[l];
      const float* vs_lo = w.vs_lo + B * m->split_off[l];
      TcGemmProblem p;
      p.M = m->M; p.N = Ld.out; p.K = Ld.in; p.batch = B;
      p.A1.hi = m->A_hi[l]; p.A1.lo = m->A_lo[l]; p.A1.ld = m->A_ld[l]; p.A1.sz = m->M * m->A_ld[l]; p.A1.major_k = 1;
      p.a_batched = 0;
      p.B1.hi = vs_hi; p.B1.lo = vs_lo; p.B1.ld = ldw; p.B1.sz = (int64_t)Ld.in * ldw; p.B1.major_k = 0;
      if (inplace_ok(m, l, V, w)) {
        // the probe's [in, out] block where the caller put it: row stride `out`, probe stride ldv; its lo part is zero by contract
        // (lo_nz[l] stays 0 after the memset above, so the kernels never touch the lo map, which only has to be encodable)
        p.B1.hi = V + Ld.woff; p.B1.lo = V + Ld.woff; p.B1.ld = Ld.out; p.B1.sz = w.ldv;
      }
      p.b_batched = 1;
      // exactly-TF32 probes (Rademacher +-1, one-hot): the split found no lo part -> skip its loads / MMAs.  Not in the
      // probe-chunked overlap mode, where a later chunk's split may still be running when the first GEMM starts.
      if (!(overlap && l == first_tc && l == 0)) p.B1.lo_nz = m->lo_nz + l;
      if (prev_hi) {
        p.A2.hi = prev_hi; p.A2.lo = prev_lo; p.A2.ld = prev_ld; p.A2.sz = m->M * (int64_t)prev_ld; p.A2.major_k = 1;
        p.a2_batched = 1;
        p.B2.hi = m->W_hi[l]; p.B2.lo = m->W_lo[l]; p.B2.ld = ldw; p.B2.sz = (int64_t)Ld.in * ldw; p.B2.major_k = 0;
        p.b2_batched = 0;
        p.K2 = Ld.in;
      }
      p.C = out_hi; p.C_lo = out_lo; p.c_sz = m->M * (int64_t)out_ld; p.c_sm = out_ld;
      p.epi.bias = V + Ld.boff; p.epi.bias_sz = w.ldv;
      if (!last) { p.epi.mask = m->dphi[l]; p.epi.mask_sm = Ld.out; }
      if (has_pre(l)) {
        // the chain link alone: T_l W_l + bias, masked, + the side stream's mask * (A_l dW_l)
        p.A1 = p.A2; p.a_batched = 1;
        p.B1 = p.B2; p.b_batched = 0; p.B1.lo_nz = nullptr;
        p.A2 = TcOperand(); p.B2 = TcOperand(); p.K2 = 0; p.a2_batched = 0; p.b2_batched = 0;
        p.epi.add = w.pre + l * pre_stride; p.epi.add_sz = m->M * (int64_t)out_ld; p.epi.add_scale = 1.f;
        LIP_CHECK_CUDA(cudaStreamWaitEvent(st, m->ev_split[l * lip_model::SPLIT_CHUNKS + 2], 0));
      }
      int rc = LIP_OK;
      if (overlap && l == first_tc && l == 0) {
        // probe-chunked: chunk c starts as soon as its split has landed
        for (int c = 0; c < NCH && !rc; ++c) {
          const int64_t b0 = c * chunk, b1 = b0 + chunk < B ? b0 + chunk : B;
          LIP_CHECK_CUDA(cudaStreamWaitEvent(st, m->ev_split[l * NCH + c], 0));
          if (b1 <= b0) continue;
          TcGemmProblem q = p;
          q.batch = b1 - b0;
          q.B1.hi += b0 * q.B1.sz; q.B1.lo += b0 * q.B1.sz;
          q.C += b0 * q.c_sz; if (q.C_lo) q.C_lo += b0 * q.c_sz;
          q.epi.bias += b0 * q.epi.bias_sz;
          rc = gemm_tc(q, st);
        }
      } else {
        if (overlap) {
          const int nch = (l == first_tc) ? NCH : 1;
          for (int c = 0; c < nch; ++c) LIP_CHECK_CUDA(cudaStreamWaitEvent(st, m->ev_split[l * NCH + c], 0));
        }
        rc = gemm_tc(p, st);
      }
      if (rc) return rc;
    } else {
      GemmProblem p;
      p.M = m->M; p.N = Ld.out; p.K = Ld.in; p.batch = B;
      p.A1 = {m->A[l], 0, Ld.in, 1};
      p.B1 = {V + Ld.woff, w.ldv, Ld.out, 1};
      if (prev_hi) {
        // a SIMT layer always reads a plain fp32 intermediate (its producer saw next_tc == false)
        p.A2 = {prev_hi, m->M * (int64_t)prev_ld, prev_ld, 1};
        p.B2 = {m->theta + Ld.woff, 0, Ld.out, 1};
        p.K2 = Ld.in;
      }
      p.C = out_hi; p.c_sz = m->M * (int64_t)out_ld; p.c_sm = out_ld;
      p.epi.C_lo = out_lo;
      p.epi.bias = V + Ld.boff; p.epi.bias_sz = w.ldv;
      if (!last) { p.epi.mask = m->dphi[l]; p.epi.mask_sm = Ld.out; }
      int rc = gemm_simt(p, st);
      if (rc) return rc;
    }
    prev_hi = out_hi; prev_lo = out_lo; prev_ld = out_ld;
  }
  return LIP_OK;
}

// ---- VJP sweep: Delta_L (plain fp32, [B,M,K] contiguous) in w.hi[src]; writes out[B,D] = scale * J^T delta + add_scale * add.
// start_layer (default nL - 1): the sweep starts at that layer with its delta already in w.hi[src] (/ w.lo[src] when that layer is a
// tensor-core layer): the fused head kernel has produced the delta of layer nL - 2 and the head's own gradients.
int vjp_sweep(lip_model* m, int src, int64_t B, const Workspace& w, float* out, float scale, const float* add,
              float add_scale, cudaStream_t st, int start_layer = -1, float* cot_in = nullptr) {
  const int nL = (int)m->L.size();
  const int lfirst = start_layer < 0 ? nL - 1 : start_layer;
  int cur = src;
  int cur_ld = m->K;
  bool cur_split = false;
  bool cur_colsum = false;   // w.colsum holds the column sums of the current delta (written by the GEMM that produced it)
  const int64_t nslots = colsum_slots(m);
  if (lfirst < nL - 1) {
    cur_ld = ld_of(m, m->L[lfirst].out);
    cur_split = m->tc_on && m->tc_layer[lfirst];
  } else if (m->tc_on && m->tc_layer[nL - 1]) {
    // the top layer runs on the tensor cores: re-lay Delta_L as padded hi/lo
    const int ldp = pad4(m->K);
    int rc = tf32_split(w.hi[cur], m->K, w.hi[cur ^ 1], w.lo[cur ^ 1], ldp, B * m->M, m->K, st);
    if (rc) return rc;
    cur ^= 1; cur_ld = ldp; cur_split = true;
  }
  // One or two probes (the mat-vecs of a single-probe Krylov recurrence): every GEMM is a handful of CTAs and the sweep is a chain of
  // launch latencies.  The weight-gradient GEMM of a layer does not feed the delta chain, so it runs on the side stream next to the
  // delta back-propagation of the same layer (fork / join with events: capturable).  The only shared buffer is the delta itself: the
  // back-propagation of layer l - 1 overwrites the buffer that the weight gradient of layer l + 1 ... l reads, hence the wait below.
  static const int fork_env = getenv("LIP_VJP_FORK") ? atoi(getenv("LIP_VJP_FORK")) : 1;
  const bool fork = fork_env && m->tc_on && m->side != nullptr && B <= 2 &&
                    (int)m->ev_split.size() >= nL * lip_model::SPLIT_CHUNKS && lip_model::SPLIT_CHUNKS >= 2;
  cudaEvent_t wg_done = nullptr;       // weight gradient still reading the delta buffer that the next back-propagation overwrites
  cudaEvent_t wg_last = nullptr;
  for (int l = lfirst; l >= 0; --l) {
    const DenseLayer& Ld = m->L[l];
    const bool tc = m->tc_on && m->tc_layer[l];
    const float* d_hi = w.hi[cur];
    const float* d_lo = cur_split ? w.lo[cur] : nullptr;
    cudaEvent_t wg_prev = wg_done;
    wg_done = nullptr;
    if (tc) {  // weight gradient [in x out] = A_l^T [in x M] * Delta [M x out]
      TcGemmProblem p;
      p.M = Ld.in; p.N = Ld.out; p.K = m->M; p.batch = B;
      p.A1.hi = m->A_hi[l]; p.A1.lo = m->A_lo[l]; p.A1.ld = m->A_ld[l]; p.A1.sz = m->M * m->A_ld[l]; p.A1.major_k = 0;
      p.a_batched = 0;
      p.B1.hi = d_hi; p.B1.lo = d_lo; p.B1.ld = cur_ld; p.B1.sz = m->M * (int64_t)cur_ld; p.B1.major_k = 0;
      p.b_batched = 1;
      p.C = out + Ld.woff; p.c_sz = w.ldo; p.c_sm = Ld.out;
      p.epi.scale = scale;
      if (add) { p.epi.add = add + Ld.woff; p.epi.add_sz = w.lda; p.epi.add_scale = add_scale; }
      if (fork) {
        cudaEvent_t ready = m->ev_split[l * lip_model::SPLIT_CHUNKS], done = m->ev_split[l * lip_model::SPLIT_CHUNKS + 1];
        LIP_CHECK_CUDA(cudaEventRecord(ready, st));
        LIP_CHECK_CUDA(cudaStreamWaitEvent(m->side, ready, 0));
        int rc = gemm_tc(p, m->side);
        if (rc) return rc;
        LIP_CHECK_CUDA(cudaEventRecord(done, m->side));
        wg_done = wg_last = done;
      } else {
        int rc = gemm_tc(p, st);
        if (rc) return rc;
      }
    } else {
      GemmProblem p;
      p.M = Ld.in; p.N = Ld.out; p.K = m->M; p.batch = B;
      p.A1 = {m->A[l], 0, 1, Ld.in};
      p.B1 = {d_hi, m->M * (int64_t)cur_ld, cur_ld, 1};
      p.C = out + Ld.woff; p.c_sz = w.ldo; p.c_sm = Ld.out;
      p.epi.scale = scale;
      if (add) { p.epi.add = add + Ld.woff; p.epi.add_sz = w.lda; p.epi.add_scale = add_scale; }
      int rc = gemm_simt(p, st);
      if (rc) return rc;
    }
    {  // bias gradient: column sums of Delta_l (from the producing GEMM's epilogue when it was a tcgen05 GEMM)
      int rcb;
      if (cur_colsum)
        rcb = launch_bias_grad(w.colsum, nullptr, nslots, Ld.out, cur_ld, B, out + Ld.boff, w.ldo, scale, add ? add + Ld.boff : nullptr,
                               w.lda, add_scale, st);
      else
        rcb = launch_bias_grad(d_hi, d_lo, m->M, Ld.out, cur_ld, B, out + Ld.boff, w.ldo, scale, add ? add + Ld.boff : nullptr, w.lda,
                               add_scale, st);
      if (rcb) return rcb;
    }
    if (l > 0) {  // Delta_{l-1} = (Delta_l W_l^T) * phi'_{l-1}
      if (wg_prev) LIP_CHECK_CUDA(cudaStreamWaitEvent(st, wg_prev, 0));     // the weight gradient of layer l + 1 has read w.hi[cur ^ 1]
      const bool next_split = m->tc_on && m->tc_layer[l - 1];
      const int nxt = cur ^ 1;
      const int nxt_ld = ld_of(m, Ld.in);
      if (tc) {
        TcGemmProblem p;
        p.M = m->M; p.N = Ld.in; p.K = Ld.out; p.batch = B;
        p.A1.hi = d_hi; p.A1.lo = d_lo; p.A1.ld = cur_ld; p.A1.sz = m->M * (int64_t)cur_ld; p.A1.major_k = 1;
        p.a_batched = 1;
        p.B1.hi = m->W_hi[l]; p.B1.lo = m->W_lo[l]; p.B1.ld = m->W_ld[l]; p.B1.sz = (int64_t)Ld.in * m->W_ld[l];
        p.B1.major_k = 1; p.b_batched = 0;
        p.C = w.hi[nxt]; p.C_lo = next_split ? w.lo[nxt] : nullptr; p.c_sz = m->M * (int64_t)nxt_ld; p.c_sm = nxt_ld;
        p.epi.mask = m->dphi[l - 1]; p.epi.mask_sm = Ld.in;
        p.colsum = w.colsum; p.colsum_ld = nxt_ld; p.colsum_sz = nslots * (int64_t)nxt_ld;
        int rc = gemm_tc(p, st);
        if (rc) return rc;
      } else if (Ld.out <= 16 && cur_ld == Ld.out && !cur_split && (size_t)Ld.in * Ld.out * sizeof(float) <= 48 * 1024) {
        const long long rows = B * m->M, total = rows * Ld.in;
        long long gb = (total + 255) / 256;
        if (gb > 148 * 16) gb = 148 * 16;
        head_dgrad_kernel<<<(unsigned)gb, 256, (size_t)Ld.in * Ld.out * sizeof(float), st>>>(
            d_hi, m->theta + Ld.woff, m->dphi[l - 1], w.hi[nxt], next_split ? w.lo[nxt] : nullptr, rows, m->M, Ld.in, Ld.out,
            nxt_ld);
        LIP_LAUNCH_CHECK();
      } else {
        GemmProblem p;
        p.M = m->M; p.N = Ld.in; p.K = Ld.out; p.batch = B;
        p.A1 = {d_hi, m->M * (int64_t)cur_ld, cur_ld, 1};
        p.B1 = {m->theta + Ld.woff, 0, 1, Ld.out};
        p.C = w.hi[nxt]; p.c_sz = m->M * (int64_t)nxt_ld; p.c_sm = nxt_ld;
        p.epi.C_lo = next_split ? w.lo[nxt] : nullptr;
        p.epi.mask = m->dphi[l - 1]; p.epi.mask_sm = Ld.in;
        int rc = gemm_simt(p, st);
        if (rc) return rc;
      }
      cur = nxt; cur_ld = nxt_ld; cur_split = next_split; cur_colsum = tc;
    } else {
      if (wg_prev) LIP_CHECK_CUDA(cudaStreamWaitEvent(st, wg_prev, 0));
      if (cot_in) {   // cotangent of the program's input: Delta_0 W_0^T, plain fp32 [B, M, in_0] (no activation below a program input)
        if (tc) {
          TcGemmProblem p;
          p.M = m->M; p.N = Ld.in; p.K = Ld.out; p.batch = B;
          p.A1.hi = d_hi; p.A1.lo = d_lo; p.A1.ld = cur_ld; p.A1.sz = m->M * (int64_t)cur_ld; p.A1.major_k = 1;
          p.a_batched = 1;
          p.B1.hi = m->W_hi[l]; p.B1.lo = m->W_lo[l]; p.B1.ld = m->W_ld[l]; p.B1.sz = (int64_t)Ld.in * m->W_ld[l];
          p.B1.major_k = 1; p.b_batched = 0;
          p.C = cot_in; p.C_lo = nullptr; p.c_sz = m->M * (int64_t)Ld.in; p.c_sm = Ld.in;
          int rc = gemm_tc(p, st);
          if (rc) return rc;
        } else {
          GemmProblem p;
          p.M = m->M; p.N = Ld.in; p.K = Ld.out; p.batch = B;
          p.A1 = {d_hi, m->M * (int64_t)cur_ld, cur_ld, 1};
          p.B1 = {m->theta + Ld.woff, 0, 1, Ld.out};
          p.C = cot_in; p.c_sz = m->M * (int64_t)Ld.in; p.c_sm = Ld.in;
          int rc = gemm_simt(p, st);
          if (rc) return rc;
        }
      }
    }
  }
  if (wg_last) LIP_CHECK_CUDA(cudaStreamWaitEvent(st, wg_last, 0));         // join: the side stream's last weight gradient
  return LIP_OK;
}

}  // namespace

namespace lip {
// ---- a dense program as the tail of a conv stage program (lip_cnn.cu) ----
lip_model* mlp_make_tail(const std::vector<ConvStage>& stages, int first, int model_type, int64_t D) {
  lip_model* t = new (std::nothrow) lip_model();
  if (!t) return nullptr;
  t->model_type = model_type;
  t->D = D;
  for (size_t i = (size_t)first; i < stages.size(); ++i) {
    const ConvStage& s = stages[i];
    DenseLayer L;
    L.in = s.cin; L.out = s.cout; L.boff = s.boff; L.woff = s.woff; L.act = s.act;
    t->L.push_back(L);
  }
  t->K = t->L.back().out;
  t->maxw = 0;
  for (auto& L : t->L) t->maxw = L.out > t->maxw ? L.out : t->maxw;
  return t;
}

static inline size_t tail_t0_floats(const lip_model* t, int64_t B) { return align_up((size_t)B * t->M * (size_t)pad4(t->L[0].in), 64); }

size_t mlp_tail_ws_bytes(const lip_model* t, int64_t B) {
  return align_up(ws_bytes(t, B), 256) + 2 * tail_t0_floats(t, B) * sizeof(float) + 512;
}

int mlp_tail_jvp(lip_model* t, const float* V, int64_t ldv, const float* T0, int64_t B, void* ws, size_t bytes, float* dlogits,
                 cudaStream_t st) {
  LIP_REQUIRE(ws && bytes >= mlp_tail_ws_bytes(t, B), "dense tail: workspace too small");
  Workspace w;
  const size_t inner = align_up(ws_bytes(t, B), 256);
  int rc = carve(t, B, ws, inner, &w);
  if (rc) return rc;
  w.ldv = ldv;
  const int in0 = t->L[0].in;
  const float* t_hi = T0;
  const float* t_lo = nullptr;
  int t_ld = in0;
  if (t->tc_on && t->tc_layer[0]) {      // layer 0 reads its input tangent as a TF32 (hi, lo) pair through TMA
    float* s_hi = (float*)align_up((uintptr_t)ws + inner, 256);
    float* s_lo = s_hi + tail_t0_floats(t, B);
    const int ld0 = pad4(in0);
    rc = tf32_split3(T0, t->M * (int64_t)in0, in0, s_hi, s_lo, t->M * (int64_t)ld0, ld0, B, t->M, in0, st, nullptr);
    if (rc) return rc;
    t_hi = s_hi; t_lo = s_lo; t_ld = ld0;
  }
  return jvp_sweep(t, V, B, w, dlogits, st, nullptr, nullptr, -1, t_hi, t_lo, t_ld);
}

int mlp_tail_vjp(lip_model* t, const float* dl, int64_t B, void* ws, size_t bytes, float* out, int64_t ldo, float scale,
                 const float* add, int64_t lda, float add_scale, float* cot_in, cudaStream_t st) {
  LIP_REQUIRE(ws && bytes >= mlp_tail_ws_bytes(t, B), "dense tail: workspace too small");
  Workspace w;
  int rc = carve(t, B, ws, align_up(ws_bytes(t, B), 256), &w);
  if (rc) return rc;
  w.ldo = ldo; w.lda = lda;
  rc = launch_scale_copy(dl, w.hi[0], B * t->M * (int64_t)t->K, 1.f, st);
  if (rc) return rc;
  return vjp_sweep(t, 0, B, w, out, scale, add, add_scale, st, -1, cot_in);
}

// ---- MLP sweep pieces shared with lip_zgrad.cu ----
size_t mlp_ws_bytes(const lip_model* m, int64_t B) { return ws_bytes(m, B); }
int mlp_ld(const lip_model* m, int width) { return ld_of(m, width); }
// forward tangent sweep keeping every hidden layer's masked tangent; reports where the probes' split weight blocks live
int mlp_jvp_keep(lip_model* m, const float* V, int64_t B, void* ws, size_t bytes, float* dl, float* const* keep_hi,
                 float* const* keep_lo, const float** vs_hi, const float** vs_lo, cudaStream_t st) {
  Workspace w;
  int rc = carve(m, B, ws, bytes, &w);
  if (rc) return rc;
  if (vs_hi) *vs_hi = w.vs_hi;
  if (vs_lo) *vs_lo = w.vs_lo;
  return jvp_sweep(m, V, B, w, dl, st, keep_hi, keep_lo);
}
}  // namespace lip

// ============================================================================================================
// C ABI
// ============================================================================================================
extern "C" {

int lip_model_create(const lip_layer_desc* layers, int32_t n_layers, int32_t model_type, int64_t num_params,
                     lip_model** out) {
  LIP_REQUIRE(layers && out && n_layers > 0, "lip_model_create: null/empty layer program");
  LIP_REQUIRE(model_type == LIP_REGRESSOR || model_type == LIP_CLASSIFIER, "lip_model_create: bad model_type %d",
              model_type);
  lip_model* m = new (std::nothrow) lip_model();
  LIP_REQUIRE(m != nullptr, "lip_model_create: out of host memory");
  m->model_type = model_type;
  m->D = num_params;
  if (layers[0].op == LIP_OP_INPUT) {   // conv stage program (lip_cnn.cu) or residual conv program (lip_resnet.cu)
    bool residual = false;
    for (int i = 0; i < n_layers; ++i) residual = residual || layers[i].op >= LIP_OP_BATCHNORM;
    int rc = residual ? resnet_parse(m, layers, n_layers, num_params) : cnn_parse(m, layers, n_layers, num_params);
    if (rc) { delete m; return rc; }
    if (model_type == LIP_REGRESSOR && m->K != 1) {
      delete m;
      set_error("lip_model_create: regressors must have one output");
      return LIP_ERR_INVALID;
    }
    *out = m;
    return LIP_OK;
  }
  int64_t counted = 0;
  for (int i = 0; i < n_layers; ++i) {
    const lip_layer_desc& d = layers[i];
    if (d.op == LIP_OP_DENSE) {
      if (d.in_features <= 0 || d.out_features <= 0 || d.bias_offset < 0 || d.kernel_offset < 0 ||
          d.bias_offset + d.out_features > num_params ||
          d.kernel_offset + (int64_t)d.in_features * d.out_features > num_params) {
        delete m;
        set_error("lip_model_create: layer %d has an invalid shape/offset", i);
        return LIP_ERR_INVALID;
      }
      if (!m->L.empty() && m->L.back().out != d.in_features) {
        set_error("lip_model_create: layer %d in_features %d != previous out_features %d", i, d.in_features,
                  m->L.back().out);
        delete m;
        return LIP_ERR_INVALID;
      }
      DenseLayer L;
      L.in = d.in_features; L.out = d.out_features; L.boff = d.bias_offset; L.woff = d.kernel_offset;
      m->L.push_back(L);
      counted += (int64_t)d.in_features * d.out_features + d.out_features;
    } else if (d.op == LIP_OP_TANH || d.op == LIP_OP_GELU_TANH || d.op == LIP_OP_RELU) {
      if (m->L.empty() || m->L.back().act != -1) {
        delete m;
        set_error("lip_model_create: activation at position %d must follow a dense layer", i);
        return LIP_ERR_INVALID;
      }
      m->L.back().act = d.op;
    } else {
      delete m;
      set_error("lip_model_create: unsupported op %d at position %d (this build executes DENSE + activations)",
                d.op, i);
      return LIP_ERR_INVALID;
    }
  }
  if (m->L.empty() || m->L.back().act != -1) {
    delete m;
    set_error("lip_model_create: the program must end with a DENSE layer");
    return LIP_ERR_INVALID;
  }
  if (counted != num_params) {
    delete m;
    set_error("lip_model_create: layers hold %lld parameters but num_params = %lld", (long long)counted,
              (long long)num_params);
    return LIP_ERR_INVALID;
  }
  m->K = m->L.back().out;
  if (model_type == LIP_REGRESSOR && m->K != 1) {
    delete m;
    set_error("lip_model_create: regressors must have one output (got %d)", m->K);
    return LIP_ERR_INVALID;
  }
  m->maxw = 0;
  for (auto& L : m->L) m->maxw = L.out > m->maxw ? L.out : m->maxw;
  *out = m;
  return LIP_OK;
}

int lip_model_destroy(lip_model* m) {
  if (!m) return LIP_OK;
  if (m->tail) { lip_model_destroy(m->tail); m->tail = nullptr; }
  m->free_cache();
  for (auto e : m->ev_split) cudaEventDestroy(e);
  if (m->ev_fork) cudaEventDestroy(m->ev_fork);
  if (m->side) cudaStreamDestroy(m->side);
  if (m->rn_stats) cudaFree(m->rn_stats);
  if (m->lo_nz) cudaFree(m->lo_nz);
  delete m;
  return LIP_OK;
}

int64_t lip_model_num_params(const lip_model* m) { return m ? m->D : -1; }
int64_t lip_model_num_outputs(const lip_model* m) { return m ? m->K : -1; }
int64_t lip_model_num_points(const lip_model* m) { return (m && m->bound) ? m->M : -1; }

int lip_model_tensor_layers(const lip_model* m) {
  if (m && m->bound && m->is_cnn && m->tail_on) return lip_model_tensor_layers(m->tail);
  if (!m || !m->bound || !m->tc_on) return 0;
  int n = 0;
  for (char c : m->tc_layer) n += c ? 1 : 0;
  return n;
}

int lip_model_fused_stages(const lip_model* m) {
  if (!m || !m->bound || !m->is_cnn) return 0;
  int n = 0;
  for (int i = 0; i + 1 < (int)m->CS.size(); ++i) n += lip::cnn_stage_fusable(m, i) ? 1 : 0;
  return n;
}

int lip_model_set_tensor_path(lip_model* m, int32_t enable) {
  LIP_REQUIRE(m != nullptr, "null model");
  m->use_tc = enable ? 1 : 0;
  return LIP_OK;
}

int lip_model_bind(lip_model* m, const float* theta, const float* Z, int64_t M, float logvar,
                   lip_stream_t stream) {
  LIP_REQUIRE(m && theta && Z && M > 0, "lip_model_bind: null argument or M <= 0");
  cudaStream_t st = (cudaStream_t)stream;
  m->free_cache();
  m->M = M;
  m->theta = theta;
  m->logvar = logvar;
  if (m->is_resnet) return resnet_bind(m, theta, Z, M, st);
  if (m->is_cnn) return cnn_bind(m, theta, Z, M, st);
  const int nL = (int)m->L.size();
  m->A.assign(nL, nullptr);
  m->dphi.assign(nL > 1 ? nL - 1 : 0, nullptr);
  for (int l = 0; l < nL; ++l) {
    LIP_CHECK_CUDA(cudaMalloc(&m->A[l], sizeof(float) * (size_t)M * m->L[l].in + 256));
    if (l < nL - 1) LIP_CHECK_CUDA(cudaMalloc(&m->dphi[l], sizeof(float) * (size_t)M * m->L[l].out + 256));
  }
  LIP_CHECK_CUDA(cudaMalloc(&m->logits, sizeof(float) * (size_t)M * m->K));
  LIP_CHECK_CUDA(cudaMalloc(&m->P, sizeof(float) * (size_t)M * m->K));
  LIP_CHECK_CUDA(cudaMalloc(&m->S, sizeof(float) * (size_t)M * m->K));
  LIP_CHECK_CUDA(cudaMemcpyAsync(m->A[0], Z, sizeof(float) * (size_t)M * m->L[0].in, cudaMemcpyDeviceToDevice, st));
  for (int l = 0; l < nL; ++l) {
    const DenseLayer& Ld = m->L[l];
    const bool last = (l == nL - 1);
    GemmProblem p;
    p.M = M; p.N = Ld.out; p.K = Ld.in; p.batch = 1;
    p.A1 = {m->A[l], 0, Ld.in, 1};
    p.B1 = {theta + Ld.woff, 0, Ld.out, 1};
    p.C = last ? m->logits : m->A[l + 1]; p.c_sz = 0; p.c_sm = Ld.out;
    p.epi.bias = theta + Ld.boff; p.epi.bias_sz = 0;
    if (!last) {
      LIP_REQUIRE(Ld.act >= 0, "lip_model_bind: hidden layer %d has no activation", l);
      p.epi.act = Ld.act;
      p.epi.dphi_out = m->dphi[l];
    }
    int rc = gemm_simt(p, st);
    if (rc) return rc;
  }
  if (m->model_type == LIP_CLASSIFIER) {
    softmax_rows_kernel<<<(unsigned)ceil_div(M, 128), 128, 0, st>>>(m->logits, m->P, m->S, M, m->K);
    LIP_LAUNCH_CHECK();
  }
  // ---- tensor-core path: decide per layer, split the shared operands once ----
  m->tc_on = false;
  m->tc_layer.assign(nL, 0);
  m->ldmax = 0;
  for (auto& Ld : m->L) m->ldmax = pad4(Ld.out) > m->ldmax ? pad4(Ld.out) : m->ldmax;
  if (m->use_tc != 0) {
    const bool avail = tc_available();
    if (m->use_tc == 1 && !avail) {
      set_error("lip_model_bind: tensor path requested but the device is not sm_100 / TMA encode unavailable");
      return LIP_ERR_UNSUPPORTED;
    }
    if (avail) {
      // fewest bound points for which the 128-row MMA tiles still beat the fp32 SIMT kernels: measured 3x faster than SIMT
      // down to M = 16 at the MNIST-MLP widths (tools/tc_small_m.py, profiles/r01_tc_small_m.txt); LIP_TC_MIN_M overrides
      static const int64_t min_m = getenv("LIP_TC_MIN_M") ? atoll(getenv("LIP_TC_MIN_M")) : 16;
      for (int l = 0; l < nL; ++l)
        m->tc_layer[l] = (m->L[l].in >= 64 && m->L[l].out >= 64 && M >= min_m && (m->L[l].out % 4 == 0)) ? 1 : 0;
      for (int l = 0; l < nL; ++l) m->tc_on = m->tc_on || m->tc_layer[l];
    }
  }
  if (m->tc_on) {
    m->A_hi.assign(nL, nullptr); m->A_lo.assign(nL, nullptr); m->W_hi.assign(nL, nullptr); m->W_lo.assign(nL, nullptr);
    m->A_ld.assign(nL, 0); m->W_ld.assign(nL, 0);
    m->max_split = 0;
    for (int l = 0; l < nL; ++l) {
      if (!m->tc_layer[l]) continue;
      const DenseLayer& Ld = m->L[l];
      const int64_t lda = pad4(Ld.in), ldw = pad4(Ld.out);
      m->A_ld[l] = lda; m->W_ld[l] = ldw;
      LIP_CHECK_CUDA(cudaMalloc(&m->A_hi[l], sizeof(float) * (size_t)M * lda + 256));
      LIP_CHECK_CUDA(cudaMalloc(&m->A_lo[l], sizeof(float) * (size_t)M * lda + 256));
      LIP_CHECK_CUDA(cudaMalloc(&m->W_hi[l], sizeof(float) * (size_t)Ld.in * ldw + 256));
      LIP_CHECK_CUDA(cudaMalloc(&m->W_lo[l], sizeof(float) * (size_t)Ld.in * ldw + 256));
      int rc = tf32_split(m->A[l], Ld.in, m->A_hi[l], m->A_lo[l], lda, M, Ld.in, st);
      if (rc) return rc;
      rc = tf32_split(theta + Ld.woff, Ld.out, m->W_hi[l], m->W_lo[l], ldw, Ld.in, Ld.out, st);
      if (rc) return rc;
      if ((int64_t)Ld.in * ldw > m->max_split) m->max_split = (int64_t)Ld.in * ldw;
    }
    m->split_off.assign(nL, 0);
    m->sum_split = 0;
    for (int l = 0; l < nL; ++l) {
      if (!m->tc_layer[l]) continue;
      m->split_off[l] = m->sum_split;
      m->sum_split += (int64_t)m->L[l].in * m->W_ld[l];
    }
    if (!m->lo_nz) LIP_CHECK_CUDA(cudaMalloc(&m->lo_nz, sizeof(int) * (nL > 0 ? nL : 1)));
    if (!m->side) {
      LIP_CHECK_CUDA(cudaStreamCreateWithFlags(&m->side, cudaStreamNonBlocking));
      LIP_CHECK_CUDA(cudaEventCreateWithFlags(&m->ev_fork, cudaEventDisableTiming));
    }
    while ((int)m->ev_split.size() < nL * lip_model::SPLIT_CHUNKS) {
      cudaEvent_t e;
      LIP_CHECK_CUDA(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
      m->ev_split.push_back(e);
    }
  }
  {
    int rc = zgrad_prepare(m, st);
    if (rc) return rc;
  }
  m->bound = true;
  return LIP_OK;
}

int lip_model_set_bn_stats(lip_model* m, const float* stats, int64_t n, lip_stream_t stream) {
  LIP_REQUIRE(m && stats && n > 0, "lip_model_set_bn_stats: null argument");
  LIP_REQUIRE(m->is_resnet && n == m->rn_nstats, "lip_model_set_bn_stats: the program's BATCHNORM ops need %lld statistics, got %lld",
              (long long)(m ? m->rn_nstats : 0), (long long)n);
  if (!m->rn_stats) LIP_CHECK_CUDA(cudaMalloc(&m->rn_stats, sizeof(float) * (size_t)n));
  LIP_CHECK_CUDA(cudaMemcpyAsync(m->rn_stats, stats, sizeof(float) * (size_t)n, cudaMemcpyDeviceToDevice, (cudaStream_t)stream));
  return LIP_OK;
}

int lip_model_outputs(lip_model* m, float* out, lip_stream_t stream) {
  LIP_REQUIRE(m && out, "null argument");
  if (!m->bound) { set_error("model not bound"); return LIP_ERR_NOT_BOUND; }
  LIP_CHECK_CUDA(cudaMemcpyAsync(out, m->logits, sizeof(float) * (size_t)m->M * m->K, cudaMemcpyDeviceToDevice,
                                 (cudaStream_t)stream));
  return LIP_OK;
}

size_t lip_workspace_bytes(const lip_model* m, int64_t B) {
  if (!m || !m->bound || B <= 0) return 0;
  return m->is_resnet ? resnet_ws_bytes(m, B) : (m->is_cnn ? cnn_ws_bytes(m, B) : ws_bytes(m, B));
}

int lip_ggn_vp(lip_model* m, const float* V, float* out, int64_t B, float recal, float alpha, void* workspace,
               size_t workspace_bytes, lip_stream_t stream) {
  return lip_ggn_vp_ex(m, V, m ? m->D : 0, out, m ? m->D : 0, B, recal, alpha, 0, workspace, workspace_bytes, stream);
}

int lip_ggn_vp_ex(lip_model* m, const float* V, int64_t ldv, float* out, int64_t ldo, int64_t B, float recal, float alpha,
                  int32_t flags, void* workspace, size_t workspace_bytes, lip_stream_t stream) {
  LIP_REQUIRE(m && V && out && B > 0, "lip_ggn_vp: null argument or B <= 0");
  LIP_REQUIRE(V != out, "lip_ggn_vp: in-place operation is not supported");
  if (!m->bound) { set_error("lip_ggn_vp: model not bound"); return LIP_ERR_NOT_BOUND; }
  LIP_REQUIRE(ldv >= m->D && ldo >= m->D, "lip_ggn_vp: row strides (%lld, %lld) are smaller than D = %lld", (long long)ldv,
              (long long)ldo, (long long)m->D);
  cudaStream_t st = (cudaStream_t)stream;
  if (m->is_resnet || m->is_cnn) {
    LIP_REQUIRE(ldv == m->D && ldo == m->D, "lip_ggn_vp_ex: conv programs take contiguous [B, D] blocks");
    if (m->is_resnet) return resnet_ggn_vp(m, V, out, B, recal, alpha, workspace, workspace_bytes, st);
    return cnn_ggn_vp(m, V, out, B, recal, alpha, workspace, workspace_bytes, st);
  }
  Workspace w;
  int rc = carve(m, B, workspace, workspace_bytes, &w);
  if (rc) return rc;
  w.ldv = ldv; w.lda = ldv; w.ldo = ldo;
  w.exact = (flags & LIP_PROBES_EXACT_TF32) != 0;
  const int nL = (int)m->L.size();
  if (head_fusable(m)) {
    // wide layers: JVP sweep up to the head's input; head JVP + output-space Hessian + head gradients + delta of the layer below in
    // one kernel; VJP sweep from the layer below
    rc = jvp_sweep(m, V, B, w, nullptr, st, nullptr, nullptr, nL - 1);
    if (rc) return rc;
    const DenseLayer& Lh = m->L[nL - 1];
    const int s = (nL - 2) & 1;
    const bool below_tc = m->tc_on && m->tc_layer[nL - 2];
    HeadArgs a{};
    a.A = m->A[nL - 1];
    a.dA = w.hi[s]; a.dA_ld = ld_of(m, Lh.in); a.dA_sz = m->M * (long long)a.dA_ld;
    a.V = V; a.theta = m->theta; a.D = w.ldv; a.ldo = w.ldo; a.lda = w.lda; a.woff = Lh.woff; a.boff = Lh.boff; a.in = Lh.in; a.K = Lh.out; a.M = m->M;
    a.P = m->model_type == LIP_CLASSIFIER ? m->P : nullptr; a.S = m->S;
    a.mask = m->dphi[nL - 2];
    a.out = out; a.scale = recal; a.add = alpha != 0.f ? V : nullptr; a.add_scale = alpha;
    a.Dn_hi = w.hi[s ^ 1]; a.Dn_lo = below_tc ? w.lo[s ^ 1] : nullptr; a.Dn_ld = a.dA_ld; a.Dn_sz = a.dA_sz;
    const size_t smem = head_smem_bytes(Lh.in, Lh.out);
    if (smem > 48 * 1024) LIP_CHECK_CUDA(cudaFuncSetAttribute(head_fused_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    for (int64_t b0 = 0; b0 < B; b0 += 65535 * 32) {
      const int64_t bc = B - b0 < 65535 * 32 ? B - b0 : 65535 * 32;
      HeadArgs c = a;
      c.V += b0 * w.ldv; c.out += b0 * w.ldo; if (c.add) c.add += b0 * w.lda;
      c.dA += b0 * a.dA_sz; c.Dn_hi += b0 * a.Dn_sz; if (c.Dn_lo) c.Dn_lo += b0 * a.Dn_sz;
      head_fused_kernel<<<(unsigned)bc, 256, smem, st>>>(c);
      LIP_LAUNCH_CHECK();
    }
    return vjp_sweep(m, s ^ 1, B, w, out, recal, alpha != 0.f ? V : nullptr, alpha, st, nL - 2);
  }
  // dlogits land in the buffer the last hidden layer did NOT use, so the VJP can ping-pong from it
  const int src = (nL - 1) & 1;
  float* dl = w.hi[src];
  rc = jvp_sweep(m, V, B, w, dl, st);
  if (rc) return rc;
  if (m->model_type == LIP_CLASSIFIER) {
    rc = launch_factor(dl, dl, m, B, 0, 1.f, st);
    if (rc) return rc;
  }
  return vjp_sweep(m, src, B, w, out, recal, alpha != 0.f ? V : nullptr, alpha, st);
}

int lip_wt_apply(lip_model* m, const float* V, float* out, int64_t B, float scale, int32_t factor,
                 void* workspace, size_t workspace_bytes, lip_stream_t stream) {
  LIP_REQUIRE(m && V && out && B > 0, "lip_wt_apply: null argument or B <= 0");
  if (!m->bound) { set_error("lip_wt_apply: model not bound"); return LIP_ERR_NOT_BOUND; }
  cudaStream_t st = (cudaStream_t)stream;
  if (m->is_resnet) return resnet_wt_apply(m, V, out, B, scale, factor, workspace, workspace_bytes, st);
  if (m->is_cnn) return cnn_wt_apply(m, V, out, B, scale, factor, workspace, workspace_bytes, st);
  Workspace w;
  int rc = carve(m, B, workspace, workspace_bytes, &w);
  if (rc) return rc;
  rc = jvp_sweep(m, V, B, w, out, st);
  if (rc) return rc;
  float s = scale;
  if (factor == LIP_FACTOR_SQRT && m->model_type == LIP_REGRESSOR) s *= expf(-0.5f * m->logvar);
  if (factor == LIP_FACTOR_SQRT && m->model_type == LIP_CLASSIFIER) return launch_factor(out, out, m, B, 1, s, st);
  if (s != 1.f) {
    int64_t n = B * m->M * m->K;
    scale_copy_kernel<<<(unsigned)ceil_div(n, 256), 256, 0, st>>>(out, out, n, s);
    LIP_LAUNCH_CHECK();
  }
  return LIP_OK;
}

int lip_w_apply(lip_model* m, const float* U, float* out, int64_t B, float scale, int32_t factor, const float* add,
                float add_scale, void* workspace, size_t workspace_bytes, lip_stream_t stream) {
  LIP_REQUIRE(m && U && out && B > 0, "lip_w_apply: null argument or B <= 0");
  if (!m->bound) { set_error("lip_w_apply: model not bound"); return LIP_ERR_NOT_BOUND; }
  cudaStream_t st = (cudaStream_t)stream;
  if (m->is_resnet) return resnet_w_apply(m, U, out, B, scale, factor, add, add_scale, workspace, workspace_bytes, st);
  if (m->is_cnn) return cnn_w_apply(m, U, out, B, scale, factor, add, add_scale, workspace, workspace_bytes, st);
  Workspace w;
  int rc = carve(m, B, workspace, workspace_bytes, &w);
  if (rc) return rc;
  float s = scale;
  if (factor == LIP_FACTOR_SQRT && m->model_type == LIP_CLASSIFIER) {
    rc = launch_factor(U, w.hi[0], m, B, 2, 1.f, st);
  } else {
    if (factor == LIP_FACTOR_SQRT) s *= expf(-0.5f * m->logvar);
    int64_t n = B * m->M * m->K;
    scale_copy_kernel<<<(unsigned)ceil_div(n, 256), 256, 0, st>>>(U, w.hi[0], n, 1.f);
    LIP_LAUNCH_CHECK();
    rc = LIP_OK;
  }
  if (rc) return rc;
  return vjp_sweep(m, 0, B, w, out, s, add, add_scale, st);
}

size_t lip_gram_workspace_bytes(const lip_model* m, int64_t block) {
  if (!m || !m->bound || block <= 0) return 0;
  size_t d = (size_t)m->M * m->K;
  return (m->is_resnet ? resnet_ws_bytes(m, block) : (m->is_cnn ? cnn_ws_bytes(m, block) : ws_bytes(m, block))) + align_up(sizeof(float) * (size_t)block * d, 256) +
         align_up(sizeof(float) * (size_t)block * (size_t)m->D, 256) + 512;
}

int lip_gram_wtw(lip_model* m, float* G, float scale, int64_t block, void* workspace, size_t workspace_bytes,
                 lip_stream_t stream) {
  LIP_REQUIRE(m && G && block > 0, "lip_gram_wtw: null argument or block <= 0");
  if (!m->bound) { set_error("lip_gram_wtw: model not bound"); return LIP_ERR_NOT_BOUND; }
  cudaStream_t st = (cudaStream_t)stream;
  const int64_t d = m->M * m->K;
  size_t need = lip_gram_workspace_bytes(m, block);
  if (workspace_bytes < need || !workspace) {
    set_error("lip_gram_wtw: workspace too small: need %zu bytes, got %zu", need, workspace_bytes);
    return LIP_ERR_WORKSPACE;
  }
  uintptr_t base = align_up((uintptr_t)workspace, 256);
  float* U = (float*)base;
  base += align_up(sizeof(float) * (size_t)block * d, 256);
  float* T = (float*)base;
  base += align_up(sizeof(float) * (size_t)block * (size_t)m->D, 256);
  void* inner = (void*)base;
  size_t inner_bytes = workspace_bytes - (base - (uintptr_t)workspace);
  for (int64_t start = 0; start < d; start += block) {
    int64_t blk = d - start < block ? d - start : block;
    onehot_rows_kernel<<<(unsigned)ceil_div(blk * d, 256), 256, 0, st>>>(U, d, start, blk);
    LIP_LAUNCH_CHECK();
    int rc = lip_w_apply(m, U, T, blk, scale, LIP_FACTOR_SQRT, nullptr, 0.f, inner, inner_bytes, stream);
    if (rc) return rc;
    rc = lip_wt_apply(m, T, G + start * d, blk, scale, LIP_FACTOR_SQRT, inner, inner_bytes, stream);
    if (rc) return rc;
  }
  symmetrize_from_lower_kernel<<<(unsigned)ceil_div(d * d, 256), 256, 0, st>>>(G, d);
  LIP_LAUNCH_CHECK();
  return LIP_OK;
}

size_t lip_gram_cross_workspace_bytes(const lip_model* mx, const lip_model* mz, int64_t block) {
  if (!mx || !mz || !mx->bound || !mz->bound || block <= 0) return 0;
  auto inner = [&](const lip_model* m) {
    return m->is_resnet ? resnet_ws_bytes(m, block) : (m->is_cnn ? cnn_ws_bytes(m, block) : ws_bytes(m, block));
  };
  size_t in_x = inner(mx), in_z = inner(mz);
  size_t dz = (size_t)mz->M * mz->K;
  return (in_x > in_z ? in_x : in_z) + align_up(sizeof(float) * (size_t)block * dz, 256) +
         align_up(sizeof(float) * (size_t)block * (size_t)mz->D, 256) + 512;
}

int lip_gram_cross(lip_model* mx, lip_model* mz, float* Gt, float scale_x, float scale_z, int64_t block, void* workspace,
                   size_t workspace_bytes, lip_stream_t stream) {
  LIP_REQUIRE(mx && mz && Gt && block > 0, "lip_gram_cross: null argument or block <= 0");
  if (!mx->bound || !mz->bound) { set_error("lip_gram_cross: model not bound"); return LIP_ERR_NOT_BOUND; }
  LIP_REQUIRE(mx->D == mz->D && mx->K == mz->K && mx->model_type == mz->model_type,
              "lip_gram_cross: the two handles must hold the same layer program (D %lld vs %lld, K %d vs %d)",
              (long long)mx->D, (long long)mz->D, mx->K, mz->K);
  cudaStream_t st = (cudaStream_t)stream;
  const int64_t dz = mz->M * mz->K, dx = mx->M * mx->K;
  size_t need = lip_gram_cross_workspace_bytes(mx, mz, block);
  if (workspace_bytes < need || !workspace) {
    set_error("lip_gram_cross: workspace too small: need %zu bytes, got %zu", need, workspace_bytes);
    return LIP_ERR_WORKSPACE;
  }
  uintptr_t base = align_up((uintptr_t)workspace, 256);
  float* U = (float*)base;
  base += align_up(sizeof(float) * (size_t)block * dz, 256);
  float* T = (float*)base;
  base += align_up(sizeof(float) * (size_t)block * (size_t)mz->D, 256);
  void* inner = (void*)base;
  size_t inner_bytes = workspace_bytes - (base - (uintptr_t)workspace);
  for (int64_t start = 0; start < dz; start += block) {
    int64_t blk = dz - start < block ? dz - start : block;
    onehot_rows_kernel<<<(unsigned)ceil_div(blk * dz, 256), 256, 0, st>>>(U, dz, start, blk);
    LIP_LAUNCH_CHECK();
    int rc = lip_w_apply(mz, U, T, blk, scale_z, LIP_FACTOR_SQRT, nullptr, 0.f, inner, inner_bytes, stream);
    if (rc) return rc;
    rc = lip_wt_apply(mx, T, Gt + start * dx, blk, scale_x, LIP_FACTOR_SQRT, inner, inner_bytes, stream);
    if (rc) return rc;
  }
  return LIP_OK;
}

}  // extern "C"
