// lip_vecops.cu — HBM-bound vector stage of CG / Lanczos / Golub-Kahan, batched over B independent columns.
//
// Reference semantics: jax.scipy.sparse.linalg.cg (call sites src/stochtrace.py:146,192, src/sample.py:71),
// matfree decomp.tridiag_sym / decomp.bidiag full re-orthogonalisation (call sites src/sample.py:114,
// src/train_inducing.py:156).  All scalars stay on the device; reductions are deterministic two-stage
// (per-CTA partials, then every consumer CTA re-sums the partials in a fixed order) — no atomics.
// Loads are float4 when pointers/strides allow, coalesced scalar otherwise.
#include "lip_common.cuh"

using namespace lip;

namespace {

constexpr int VT = 256;            // threads per CTA
constexpr int MAX_CHUNKS = 1024;   // partials per column

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
  return t;
}

// sum of partial[0..nch) in a fixed order, result broadcast to the whole CTA
__device__ __forceinline__ float sum_partials(const float* part, int nch, float* sm) {
  float v = 0.f;
  for (int i = threadIdx.x; i < nch; i += blockDim.x) v += part[i];
  return block_sum(v, sm);
}

inline int num_chunks(int64_t n) {
  int64_t c = ceil_div(n, 8192);
  if (c < 1) c = 1;
  if (c > MAX_CHUNKS) c = MAX_CHUNKS;
  return (int)c;
}

inline bool aligned16(const void* p) { return ((uintptr_t)p & 15) == 0; }

// ---- dot partials: part[b][chunk] = sum over the chunk's grid-stride elements of x*y -------------------------
template <int VEC>
__global__ void __launch_bounds__(VT) dot_partial_kernel(const float* __restrict__ x, const float* __restrict__ y,
                                                         float* __restrict__ part, int64_t n, int64_t ldx,
                                                         int64_t ldy) {
  __shared__ float sm[32];
  const int b = blockIdx.y, nch = gridDim.x;
  const float* xb = x + (int64_t)b * ldx;
  const float* yb = y + (int64_t)b * ldy;
  float acc = 0.f;
  if (VEC == 4) {
    const int64_t n4 = n >> 2;
    const float4* x4 = reinterpret_cast<const float4*>(xb);
    const float4* y4 = reinterpret_cast<const float4*>(yb);
    for (int64_t i = blockIdx.x * (int64_t)VT + threadIdx.x; i < n4; i += (int64_t)nch * VT) {
      float4 a = __ldg(x4 + i), c = __ldg(y4 + i);
      acc += (a.x * c.x + a.y * c.y) + (a.z * c.z + a.w * c.w);
    }
    if (blockIdx.x == 0 && threadIdx.x < (n & 3)) acc += xb[(n4 << 2) + threadIdx.x] * yb[(n4 << 2) + threadIdx.x];
  } else {
    for (int64_t i = blockIdx.x * (int64_t)VT + threadIdx.x; i < n; i += (int64_t)nch * VT) acc += xb[i] * yb[i];
  }
  acc = block_sum(acc, sm);
  if (threadIdx.x == 0) part[(int64_t)b * nch + blockIdx.x] = acc;
}

__global__ void reduce_partials_kernel(const float* __restrict__ part, int nch, float* __restrict__ out, int do_sqrt) {
  __shared__ float sm[32];
  const int b = blockIdx.x;
  float v = sum_partials(part + (int64_t)b * nch, nch, sm);
  if (threadIdx.x == 0) out[b] = do_sqrt ? sqrtf(v) : v;
}

int dot_partials(const float* x, const float* y, float* part, int64_t n, int64_t B, int64_t ldx, int64_t ldy,
                 cudaStream_t st, int* nch_out) {
  int nch = num_chunks(n);
  *nch_out = nch;
  dim3 grid(nch, (unsigned)B);
  bool v4 = aligned16(x) && aligned16(y) && (ldx % 4 == 0) && (ldy % 4 == 0);
  if (v4) dot_partial_kernel<4><<<grid, VT, 0, st>>>(x, y, part, n, ldx, ldy);
  else dot_partial_kernel<1><<<grid, VT, 0, st>>>(x, y, part, n, ldx, ldy);
  LIP_LAUNCH_CHECK();
  return LIP_OK;
}

// ---- elementwise ---------------------------------------------------------------------------------------------------
__global__ void axpby_kernel(const float* __restrict__ a, const float* __restrict__ x, const float* __restrict__ c,
                             float* __restrict__ y, int64_t n, int64_t ldx, int64_t ldy) {
  const int b = blockIdx.y;
  const float av = a ? a[b] : 1.f;
  const float cv = c ? c[b] : 0.f;
  const float* xb = x + (int64_t)b * ldx;
  float* yb = y + (int64_t)b * ldy;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    float v = av * xb[i];
    if (c) v += cv * yb[i];
    yb[i] = v;
  }
}

__global__ void scale_kernel(const float* __restrict__ s, int invert, const float* __restrict__ x,
                             float* __restrict__ y, int64_t n, int64_t ldx, int64_t ldy) {
  const int b = blockIdx.y;
  const float sv = invert ? 1.f / s[b] : s[b];
  const float* xb = x + (int64_t)b * ldx;
  float* yb = y + (int64_t)b * ldy;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    yb[i] = sv * xb[i];
}

// one thread per packed byte -> 8 consecutive floats; flat grid-stride over all B * ceil(n/8) bytes
__global__ void __launch_bounds__(256) unpack_rademacher_kernel(const uint8_t* __restrict__ bits, int64_t ldbits,
                                                                float* __restrict__ out, int64_t ldo, int64_t n, int64_t B) {
  const int64_t nbytes = (n + 7) >> 3, total = nbytes * B;
  for (int64_t idx = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; idx < total; idx += (int64_t)gridDim.x * blockDim.x) {
    const int64_t b = idx / nbytes, i = idx - b * nbytes;
    const unsigned v = bits[b * ldbits + i];
    float* o = out + b * ldo + (i << 3);
    const int64_t left = n - (i << 3);
#pragma unroll
    for (int k = 0; k < 8; ++k)
      if (k < left) o[k] = ((v >> (7 - k)) & 1u) ? 1.f : -1.f;
  }
}

// Coalesced form for even n (rows are then 8-byte aligned): a thread writes ONE float2 (elements 2j, 2j+1), so a warp stores 256
// contiguous bytes per instruction; four lanes share a packed byte (broadcast load).  The byte-per-thread form above makes every
// lane write 8 scalars 32 bytes apart (8 store instructions per warp, each touching 32 sectors).
__global__ void __launch_bounds__(256) unpack_rademacher2_kernel(const uint8_t* __restrict__ bits, int64_t ldbits,
                                                                 float* __restrict__ out, int64_t ldo, int64_t n, int64_t B) {
  const int64_t half = n >> 1;
  const int b = blockIdx.y;
  const uint8_t* row = bits + (int64_t)b * ldbits;
  float2* o = reinterpret_cast<float2*>(out + (int64_t)b * ldo);
  for (int64_t j = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; j < half; j += (int64_t)gridDim.x * blockDim.x) {
    const unsigned v = __ldg(row + (j >> 2));
    const int sh = 6 - 2 * (int)(j & 3);                 // element 2j is bit 7 - (2j & 7) of its byte (numpy.packbits order)
    float2 w;
    w.x = ((v >> (sh + 1)) & 1u) ? 1.f : -1.f;
    w.y = ((v >> sh) & 1u) ? 1.f : -1.f;
    o[j] = w;
  }
}

inline unsigned ew_blocks(int64_t n) {
  int64_t g = ceil_div(n, 256 * 4);
  if (g > 148 * 8) g = 148 * 8;
  if (g < 1) g = 1;
  return (unsigned)g;
}

// ---- CG ------------------------------------------------------------------------------------------------------------------
// scratch layout (floats): pAp partials [B*nch] | rr partials [B*nch] | gamma_next [B]
__global__ void __launch_bounds__(VT) cg_update_kernel(float* __restrict__ x, float* __restrict__ r,
                                                       const float* __restrict__ p, const float* __restrict__ Ap,
                                                       const float* __restrict__ gamma, const int* __restrict__ active,
                                                       const float* __restrict__ pap_part, float* __restrict__ rr_part,
                                                       int64_t n) {
  __shared__ float sm[32];
  const int b = blockIdx.y, nch = gridDim.x;
  const bool act = active[b] != 0;
  float pap = sum_partials(pap_part + (int64_t)b * nch, nch, sm);
  const float a = act ? gamma[b] / pap : 0.f;
  float* xb = x + (int64_t)b * n;
  float* rb = r + (int64_t)b * n;
  const float* pb = p + (int64_t)b * n;
  const float* apb = Ap + (int64_t)b * n;
  float acc = 0.f;
  for (int64_t i = blockIdx.x * (int64_t)VT + threadIdx.x; i < n; i += (int64_t)nch * VT) {
    float rv = rb[i];
    if (act) {
      xb[i] += a * pb[i];
      rv -= a * apb[i];
      rb[i] = rv;
    }
    acc += rv * rv;
  }
  acc = block_sum(acc, sm);
  if (threadIdx.x == 0) rr_part[(int64_t)b * nch + blockIdx.x] = acc;
}

__global__ void __launch_bounds__(VT) cg_direction_kernel(const float* __restrict__ r, float* __restrict__ p,
                                                          const float* __restrict__ gamma,
                                                          const int* __restrict__ active,
                                                          const float* __restrict__ rr_part,
                                                          float* __restrict__ gamma_next, int64_t n) {
  __shared__ float sm[32];
  const int b = blockIdx.y, nch = gridDim.x;
  float g1 = sum_partials(rr_part + (int64_t)b * nch, nch, sm);
  if (blockIdx.x == 0 && threadIdx.x == 0) gamma_next[b] = g1;
  if (active[b] == 0) return;
  const float beta = g1 / gamma[b];
  const float* rb = r + (int64_t)b * n;
  float* pb = p + (int64_t)b * n;
  for (int64_t i = blockIdx.x * (int64_t)VT + threadIdx.x; i < n; i += (int64_t)nch * VT) pb[i] = rb[i] + beta * pb[i];
}

__global__ void cg_commit_kernel(float* gamma, const float* gamma_next, const float* thresh, int* active, int* iters,
                                 int B) {
  int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B) return;
  if (active[b]) {
    float g = gamma_next[b];
    gamma[b] = g;
    iters[b] += 1;
    active[b] = (g > thresh[b]) ? 1 : 0;
  }
}

__global__ void cg_init_scalars_kernel(const float* bb, float* gamma, float* thresh, int* active, int* iters, float tol,
                                       float atol, int B) {
  int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B) return;
  float g = bb[b];
  float t = fmaxf(tol * tol * g, atol * atol);
  gamma[b] = g;
  thresh[b] = t;
  active[b] = (g > t) ? 1 : 0;
  iters[b] = 0;
}

// ---- re-orthogonalisation ---------------------------------------------------------------------------------------------
constexpr int RCHUNK = 2048;  // n-elements per CTA in the projection kernel (8 KB of w in shared memory)

// part[b][chunk][j] = sum_{i in chunk} Q[b][j][i] * w[b][i]   for j < kk
template <int VEC>
__global__ void __launch_bounds__(VT) reorth_project_kernel(const float* __restrict__ Q, int64_t ldq, int64_t kmax,
                                                            int kk, const float* __restrict__ w, int64_t ldw,
                                                            float* __restrict__ part, int64_t n) {
  __shared__ __align__(16) float ws[RCHUNK];
  const int b = blockIdx.y, chunk = blockIdx.x, nch = gridDim.x;
  const int64_t i0 = (int64_t)chunk * RCHUNK;
  const int len = (int)((n - i0) < RCHUNK ? (n - i0) : RCHUNK);
  const float* wb = w + (int64_t)b * ldw + i0;
  for (int i = threadIdx.x; i < RCHUNK; i += VT) ws[i] = (i < len) ? wb[i] : 0.f;
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = VT >> 5;
  const float* Qb = Q + (int64_t)b * kmax * ldq + i0;
  float* pb = part + ((int64_t)b * nch + chunk) * kmax;
  for (int j = warp; j < kk; j += nw) {
    const float* q = Qb + (int64_t)j * ldq;
    float acc = 0.f;
    if (VEC == 4) {
      const float4* q4 = reinterpret_cast<const float4*>(q);
      const float4* w4 = reinterpret_cast<const float4*>(ws);
      const int len4 = len >> 2;
#pragma unroll 4
      for (int i = lane; i < len4; i += 32) {
        float4 a = __ldg(q4 + i), c = w4[i];
        acc += (a.x * c.x + a.y * c.y) + (a.z * c.z + a.w * c.w);
      }
      for (int i = (len4 << 2) + lane; i < len; i += 32) acc += q[i] * ws[i];
    } else {
#pragma unroll 4
      for (int i = lane; i < len; i += 32) acc += q[i] * ws[i];
    }
    acc = warp_sum(acc);
    if (lane == 0) pb[j] = acc;
  }
}

// h[b][j] = sum_chunk part[b][chunk][j];  optionally also h_out[b][j] = h[b][j].
// Block = 32 coefficients x 8 chunk lanes: lane y sums chunks y, y+8, ... (4 loads in flight), then the 8 partial sums are
// added in a fixed order (deterministic).  The one-thread-per-coefficient version cost 45 us per call (latency of a
// 733-long serial chain) against ~120 us for the projection itself.
__global__ void __launch_bounds__(256) reorth_coeff_kernel(const float* __restrict__ part, int nch, int64_t kmax, int kk,
                                                           float* __restrict__ h, float* __restrict__ h_out) {
  __shared__ float sm[8][33];
  const int b = blockIdx.y;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const int j = blockIdx.x * 32 + tx;
  float acc = 0.f;
  if (j < kk) {
    const float* p = part + (int64_t)b * nch * kmax + j;
    int c = ty;
    for (; c + 24 < nch; c += 32) {
      const float a0 = p[(int64_t)c * kmax], a1 = p[(int64_t)(c + 8) * kmax];
      const float a2 = p[(int64_t)(c + 16) * kmax], a3 = p[(int64_t)(c + 24) * kmax];
      acc += (a0 + a1) + (a2 + a3);
    }
    for (; c < nch; c += 8) acc += p[(int64_t)c * kmax];
  }
  sm[ty][tx] = acc;
  __syncthreads();
  if (ty == 0 && j < kk) {
    float t = 0.f;
#pragma unroll
    for (int y = 0; y < 8; ++y) t += sm[y][tx];
    h[(int64_t)b * kmax + j] = t;
    if (h_out) h_out[(int64_t)b * kmax + j] = t;
  }
}

// out[b][i] = base[b][i] + sign * sum_{j<kk} h[b][j] Q[b][j][i];  nrm_part[b][cta] = sum out^2 (optional)
template <int VEC>
__global__ void __launch_bounds__(VT) basis_axpy_kernel(const float* __restrict__ Q, int64_t ldq, int64_t kmax, int kk,
                                                        const float* __restrict__ h, int64_t ldh, float sign,
                                                        const float* base, int64_t ldb,
                                                        float* out, int64_t ldo,
                                                        float* __restrict__ nrm_part, int64_t n) {
  extern __shared__ float hs[];
  __shared__ float sm[32];
  const int b = blockIdx.y;
  for (int j = threadIdx.x; j < kk; j += VT) hs[j] = sign * h[(int64_t)b * ldh + j];
  __syncthreads();
  const float* Qb = Q + (int64_t)b * kmax * ldq;
  const float* bb = base ? base + (int64_t)b * ldb : nullptr;
  float* ob = out + (int64_t)b * ldo;
  float nacc = 0.f;
  if (VEC == 4) {
    const int64_t n4 = n >> 2;
    for (int64_t i = blockIdx.x * (int64_t)VT + threadIdx.x; i < n4; i += (int64_t)gridDim.x * VT) {
      float4 acc = bb ? *reinterpret_cast<const float4*>(bb + (i << 2)) : make_float4(0.f, 0.f, 0.f, 0.f);
      const float4* q = reinterpret_cast<const float4*>(Qb) + i;
      const int64_t step = ldq >> 2;
      int j = 0;
      for (; j + 4 <= kk; j += 4) {
        float4 q0 = __ldg(q + (int64_t)(j + 0) * step), q1 = __ldg(q + (int64_t)(j + 1) * step);
        float4 q2 = __ldg(q + (int64_t)(j + 2) * step), q3 = __ldg(q + (int64_t)(j + 3) * step);
        float h0 = hs[j], h1 = hs[j + 1], h2 = hs[j + 2], h3 = hs[j + 3];
        acc.x += h0 * q0.x + h1 * q1.x + h2 * q2.x + h3 * q3.x;
        acc.y += h0 * q0.y + h1 * q1.y + h2 * q2.y + h3 * q3.y;
        acc.z += h0 * q0.z + h1 * q1.z + h2 * q2.z + h3 * q3.z;
        acc.w += h0 * q0.w + h1 * q1.w + h2 * q2.w + h3 * q3.w;
      }
      for (; j < kk; ++j) {
        float4 q0 = __ldg(q + (int64_t)j * step);
        float h0 = hs[j];
        acc.x += h0 * q0.x; acc.y += h0 * q0.y; acc.z += h0 * q0.z; acc.w += h0 * q0.w;
      }
      *reinterpret_cast<float4*>(ob + (i << 2)) = acc;
      nacc += (acc.x * acc.x + acc.y * acc.y) + (acc.z * acc.z + acc.w * acc.w);
    }
    if (blockIdx.x == 0 && threadIdx.x < (n & 3)) {
      int64_t i = (n4 << 2) + threadIdx.x;
      float acc = bb ? bb[i] : 0.f;
      for (int j = 0; j < kk; ++j) acc += hs[j] * Qb[(int64_t)j * ldq + i];
      ob[i] = acc;
      nacc += acc * acc;
    }
  } else {
    for (int64_t i = blockIdx.x * (int64_t)VT + threadIdx.x; i < n; i += (int64_t)gridDim.x * VT) {
      float acc = bb ? bb[i] : 0.f;
      int j = 0;
      for (; j + 4 <= kk; j += 4) {
        float q0 = Qb[(int64_t)(j + 0) * ldq + i], q1 = Qb[(int64_t)(j + 1) * ldq + i];
        float q2 = Qb[(int64_t)(j + 2) * ldq + i], q3 = Qb[(int64_t)(j + 3) * ldq + i];
        acc += hs[j] * q0 + hs[j + 1] * q1 + hs[j + 2] * q2 + hs[j + 3] * q3;
      }
      for (; j < kk; ++j) acc += hs[j] * Qb[(int64_t)j * ldq + i];
      ob[i] = acc;
      nacc += acc * acc;
    }
  }
  if (nrm_part) {
    nacc = block_sum(nacc, sm);
    if (threadIdx.x == 0) nrm_part[(int64_t)b * gridDim.x + blockIdx.x] = nacc;
  }
}

inline int axpy_blocks(int64_t n, int vec) {
  int64_t g = ceil_div(n, (int64_t)VT * vec);
  if (g > MAX_CHUNKS) g = MAX_CHUNKS;
  if (g < 1) g = 1;
  return (int)g;
}

}  // namespace

extern "C" {

size_t lip_dot_scratch_bytes(int64_t n, int64_t B) {
  return sizeof(float) * ((size_t)B * (2 * MAX_CHUNKS + 4)) + 256;
}

int lip_dot(const float* x, const float* y, float* out, int64_t n, int64_t B, int64_t ldx, int64_t ldy,
            void* scratch, lip_stream_t stream) {
  LIP_REQUIRE(B <= 65535, "lip_dot: at most 65535 columns per call (grid.y limit), got %lld: split the batch", (long long)B);
  LIP_REQUIRE(x && y && out && scratch && n > 0 && B > 0, "lip_dot: bad argument");
  cudaStream_t st = (cudaStream_t)stream;
  int nch;
  int rc = dot_partials(x, y, (float*)scratch, n, B, ldx, ldy, st, &nch);
  if (rc) return rc;
  reduce_partials_kernel<<<(unsigned)B, 256, 0, st>>>((float*)scratch, nch, out, 0);
  LIP_LAUNCH_CHECK();
  return LIP_OK;
}

int lip_axpby(const float* a, const float* x, const float* c, float* y, int64_t n, int64_t B, int64_t ldx,
              int64_t ldy, lip_stream_t stream) {
  LIP_REQUIRE(B <= 65535, "lip_axpby: at most 65535 columns per call (grid.y limit), got %lld: split the batch", (long long)B);
  LIP_REQUIRE(x && y && n > 0 && B > 0, "lip_axpby: bad argument");
  dim3 grid(ew_blocks(n), (unsigned)B);
  axpby_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(a, x, c, y, n, ldx, ldy);
  LIP_LAUNCH_CHECK();
  return LIP_OK;
}

int lip_scale(const float* s, int32_t invert, const float* x, float* y, int64_t n, int64_t B, int64_t ldx,
              int64_t ldy, lip_stream_t stream) {
  LIP_REQUIRE(B <= 65535, "lip_scale: at most 65535 columns per call (grid.y limit), got %lld: split the batch", (long long)B);
  LIP_REQUIRE(s && x && y && n > 0 && B > 0, "lip_scale: bad argument");
  dim3 grid(ew_blocks(n), (unsigned)B);
  scale_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(s, invert, x, y, n, ldx, ldy);
  LIP_LAUNCH_CHECK();
  return LIP_OK;
}

int lip_unpack_rademacher(const uint8_t* bits, int64_t ldbits, float* out, int64_t n, int64_t B, lip_stream_t stream) {
  return lip_unpack_rademacher_ld(bits, ldbits, out, n, n, B, stream);
}

int lip_unpack_rademacher_ld(const uint8_t* bits, int64_t ldbits, float* out, int64_t ldo, int64_t n, int64_t B, lip_stream_t stream) {
  LIP_REQUIRE(bits && out && n > 0 && B > 0 && ldbits * 8 >= n && ldo >= n, "lip_unpack_rademacher: bad argument");
  if (n % 2 == 0 && ldo % 2 == 0 && B <= 65535 && ((uintptr_t)out & 7) == 0) {
    int64_t gx = ceil_div(n / 2, 256 * 4);
    if (gx > 148 * 4) gx = 148 * 4;
    if (gx < 1) gx = 1;
    dim3 grid((unsigned)gx, (unsigned)B);
    unpack_rademacher2_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(bits, ldbits, out, ldo, n, B);
    LIP_LAUNCH_CHECK();
    return LIP_OK;
  }
  int64_t g = ceil_div((n + 7) / 8 * B, 256);
  if (g > 148 * 16) g = 148 * 16;
  unpack_rademacher_kernel<<<(unsigned)g, 256, 0, (cudaStream_t)stream>>>(bits, ldbits, out, ldo, n, B);
  LIP_LAUNCH_CHECK();
  return LIP_OK;
}

int lip_cg_init(const float* b, float* x, float* r, float* p, float* gamma, float* thresh, int32_t* active,
                int32_t* iters, float tol, float atol, int64_t n, int64_t B, void* scratch, lip_stream_t stream) {
  LIP_REQUIRE(B <= 65535, "lip_cg_init: at most 65535 columns per call (grid.y limit), got %lld: split the batch", (long long)B);
  LIP_REQUIRE(b && x && r && p && gamma && thresh && active && iters && scratch && n > 0 && B > 0,
              "lip_cg_init: bad argument");
  cudaStream_t st = (cudaStream_t)stream;
  LIP_CHECK_CUDA(cudaMemsetAsync(x, 0, sizeof(float) * (size_t)n * B, st));
  LIP_CHECK_CUDA(cudaMemcpyAsync(r, b, sizeof(float) * (size_t)n * B, cudaMemcpyDeviceToDevice, st));
  LIP_CHECK_CUDA(cudaMemcpyAsync(p, b, sizeof(float) * (size_t)n * B, cudaMemcpyDeviceToDevice, st));
  float* part = (float*)scratch;
  float* bb = part + (size_t)B * MAX_CHUNKS * 2;
  int nch;
  int rc = dot_partials(b, b, part, n, B, n, n, st, &nch);
  if (rc) return rc;
  reduce_partials_kernel<<<(unsigned)B, 256, 0, st>>>(part, nch, bb, 0);
  LIP_LAUNCH_CHECK();
  cg_init_scalars_kernel<<<(unsigned)ceil_div(B, 128), 128, 0, st>>>(bb, gamma, thresh, active, iters, tol, atol, (int)B);
  LIP_LAUNCH_CHECK();
  return LIP_OK;
}

int lip_cg_step(float* x, float* r, float* p, const float* Ap, float* gamma, const float* thresh, int32_t* active,
                int32_t* iters, int64_t n, int64_t B, void* scratch, lip_stream_t stream) {
  LIP_REQUIRE(B <= 65535, "lip_cg_step: at most 65535 columns per call (grid.y limit), got %lld: split the batch", (long long)B);
  LIP_REQUIRE(x && r && p && Ap && gamma && thresh && active && iters && scratch && n > 0 && B > 0,
              "lip_cg_step: bad argument");
  cudaStream_t st = (cudaStream_t)stream;
  float* pap_part = (float*)scratch;
  float* rr_part = pap_part + (size_t)B * MAX_CHUNKS;
  float* gamma_next = rr_part + (size_t)B * MAX_CHUNKS;
  int nch;
  int rc = dot_partials(p, Ap, pap_part, n, B, n, n, st, &nch);
  if (rc) return rc;
  dim3 grid(nch, (unsigned)B);
  cg_update_kernel<<<grid, VT, 0, st>>>(x, r, p, Ap, gamma, active, pap_part, rr_part, n);
  LIP_LAUNCH_CHECK();
  cg_direction_kernel<<<grid, VT, 0, st>>>(r, p, gamma, active, rr_part, gamma_next, n);
  LIP_LAUNCH_CHECK();
  cg_commit_kernel<<<(unsigned)ceil_div(B, 128), 128, 0, st>>>(gamma, gamma_next, thresh, active, iters, (int)B);
  LIP_LAUNCH_CHECK();
  return LIP_OK;
}

size_t lip_reorth_scratch_bytes(int64_t n, int64_t B, int64_t kmax) {
  size_t nch = (size_t)ceil_div(n, RCHUNK);
  return sizeof(float) * ((size_t)B * nch * (size_t)kmax + (size_t)B * kmax + (size_t)B * MAX_CHUNKS) + 1024;
}

int lip_reorth(const float* Q, int64_t ldq, int64_t kk, int64_t kmax, float* w, int64_t ldw, float* h_out,
               float* norm_out, int32_t passes, int64_t n, int64_t B, void* scratch, lip_stream_t stream) {
  LIP_REQUIRE(B <= 65535, "lip_reorth: at most 65535 columns per call (grid.y limit), got %lld: split the batch", (long long)B);
  LIP_REQUIRE(Q && w && scratch && n > 0 && B > 0 && kk >= 0 && kk <= kmax && ldq >= n && ldw >= n,
              "lip_reorth: bad argument");
  LIP_REQUIRE(passes == 1 || passes == 2, "lip_reorth: passes must be 1 or 2");
  LIP_REQUIRE(kk * sizeof(float) <= 160 * 1024, "lip_reorth: basis too deep (kk=%lld)", (long long)kk);
  cudaStream_t st = (cudaStream_t)stream;
  const int nch = (int)ceil_div(n, RCHUNK);
  float* part = (float*)scratch;
  float* h = part + (size_t)B * nch * kmax;
  float* nrm_part = h + (size_t)B * kmax;
  const bool v4 = aligned16(Q) && aligned16(w) && (ldq % 4 == 0) && (ldw % 4 == 0);
  const int nb = axpy_blocks(n, v4 ? 4 : 1);
  const size_t hs_bytes = sizeof(float) * (size_t)(kk > 0 ? kk : 1);
  if (hs_bytes > 48 * 1024) {
    LIP_CHECK_CUDA(cudaFuncSetAttribute(basis_axpy_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024));
    LIP_CHECK_CUDA(cudaFuncSetAttribute(basis_axpy_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024));
  }
  for (int pass = 0; pass < passes; ++pass) {
    const bool lastp = (pass == passes - 1);
    if (kk > 0) {
      dim3 g1(nch, (unsigned)B);
      if (v4) reorth_project_kernel<4><<<g1, VT, 0, st>>>(Q, ldq, kmax, (int)kk, w, ldw, part, n);
      else reorth_project_kernel<1><<<g1, VT, 0, st>>>(Q, ldq, kmax, (int)kk, w, ldw, part, n);
      LIP_LAUNCH_CHECK();
      dim3 g2((unsigned)ceil_div(kk, 32), (unsigned)B);
      reorth_coeff_kernel<<<g2, 256, 0, st>>>(part, nch, kmax, (int)kk, h, pass == 0 ? h_out : nullptr);
      LIP_LAUNCH_CHECK();
    }
    if (kk > 0 || (lastp && norm_out)) {
      dim3 g3(nb, (unsigned)B);
      float* np = (lastp && norm_out) ? nrm_part : nullptr;
      if (v4) basis_axpy_kernel<4><<<g3, VT, hs_bytes, st>>>(Q, ldq, kmax, (int)kk, h, kmax, -1.f, w, ldw, w, ldw, np, n);
      else basis_axpy_kernel<1><<<g3, VT, hs_bytes, st>>>(Q, ldq, kmax, (int)kk, h, kmax, -1.f, w, ldw, w, ldw, np, n);
      LIP_LAUNCH_CHECK();
    }
  }
  if (norm_out) {
    reduce_partials_kernel<<<(unsigned)B, 256, 0, st>>>(nrm_part, nb, norm_out, 1);
    LIP_LAUNCH_CHECK();
  }
  return LIP_OK;
}

int lip_basis_combine(const float* Q, int64_t ldq, int64_t kk, int64_t kmax, const float* c, int64_t ldc, float* out,
                      int64_t ldo, int64_t n, int64_t B, lip_stream_t stream) {
  LIP_REQUIRE(B <= 65535, "lip_basis_combine: at most 65535 columns per call (grid.y limit), got %lld: split the batch", (long long)B);
  LIP_REQUIRE(Q && c && out && n > 0 && B > 0 && kk > 0 && kk <= kmax && ldq >= n && ldo >= n,
              "lip_basis_combine: bad argument");
  LIP_REQUIRE(kk * sizeof(float) <= 160 * 1024, "lip_basis_combine: basis too deep");
  cudaStream_t st = (cudaStream_t)stream;
  const bool v4 = aligned16(Q) && aligned16(out) && (ldq % 4 == 0) && (ldo % 4 == 0);
  const int nb = axpy_blocks(n, v4 ? 4 : 1);
  const size_t hs_bytes = sizeof(float) * (size_t)kk;
  if (hs_bytes > 48 * 1024) {
    LIP_CHECK_CUDA(cudaFuncSetAttribute(basis_axpy_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024));
    LIP_CHECK_CUDA(cudaFuncSetAttribute(basis_axpy_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024));
  }
  dim3 g(nb, (unsigned)B);
  if (v4) basis_axpy_kernel<4><<<g, VT, hs_bytes, st>>>(Q, ldq, kmax, (int)kk, c, ldc, 1.f, nullptr, 0, out, ldo, nullptr, n);
  else basis_axpy_kernel<1><<<g, VT, hs_bytes, st>>>(Q, ldq, kmax, (int)kk, c, ldc, 1.f, nullptr, 0, out, ldo, nullptr, n);
  LIP_LAUNCH_CHECK();
  return LIP_OK;
}

}  // extern "C"
