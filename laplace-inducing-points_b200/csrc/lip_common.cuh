// lip_common.cuh — shared helpers for liblip_b200 (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdarg.h>
#include <string.h>

#include "../../include/lip_b200.h"

namespace lip {

void set_error(const char* fmt, ...);
const char* get_error();
long long launches();

#define LIP_CHECK_CUDA(expr)                                                              \
  do {                                                                                    \
    cudaError_t _e = (expr);                                                              \
    if (_e != cudaSuccess) {                                                              \
      ::lip::set_error("%s:%d CUDA error %s: %s", __FILE__, __LINE__, #expr,              \
                       cudaGetErrorString(_e));                                           \
      return LIP_ERR_CUDA;                                                                \
    }                                                                                     \
  } while (0)

#define LIP_REQUIRE(cond, ...)                                                            \
  do {                                                                                    \
    if (!(cond)) {                                                                        \
      ::lip::set_error(__VA_ARGS__);                                                      \
      return LIP_ERR_INVALID;                                                             \
    }                                                                                     \
  } while (0)

void count_launch();
// every kernel launch in the library is followed by this macro: it checks the launch and counts it
#define LIP_LAUNCH_CHECK()                    \
  do {                                        \
    ::lip::count_launch();                    \
    LIP_CHECK_CUDA(cudaGetLastError());       \
  } while (0)

static inline int64_t ceil_div(int64_t a, int64_t b) { return (a + b - 1) / b; }
static inline size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

// ---- generic strided batched GEMM (SIMT fp32) with the fused epilogue every hot-path stage needs -------
//   acc[z][m][n] = sum_k A1[z][m][k] B1[z][k][n]  (+ sum_k A2[z][m][k] B2[z][k][n])
//   v = scale*acc + bias[z][n];  (act) ;  v *= mask[m][n];  v += add_scale * add[z][m][n]
//   C[z][m][n] = v;   dphi_out[z][m][n] = act'(pre-activation)   (forward pass only)
// division by a runtime constant without the integer divider: q = (umulhi(n, mul) + n) >> shift for 0 <= n < 2^31
struct FastDiv {
  uint32_t d = 1, mul = 0, shift = 0;
  FastDiv() {}
  explicit FastDiv(uint32_t div) : d(div) {
    shift = 0;
    while ((1ull << shift) < div) ++shift;
    mul = (uint32_t)((((1ull << shift) - div) << 32) / div + 1);
  }
  __host__ __device__ __forceinline__ uint32_t div(uint32_t n) const {
#ifdef __CUDA_ARCH__
    return (uint32_t)(((uint64_t)__umulhi(n, mul) + n) >> shift);
#else
    return n / d;
#endif
  }
  __host__ __device__ __forceinline__ void divmod(uint32_t n, uint32_t& q, uint32_t& r) const { q = div(n); r = n - q * d; }
};

// Implicit-GEMM view of an NHWC image [*, Hi, Wi, C] as the A operand of a GEMM (no im2col buffer).  patch(r, kc) is the
// im2col element of output pixel r = (mz, y, x) and patch column kc = (dy, dx, c) (the row order of a flax HWIO kernel):
//   mode 1:  A[m][k] = patch(r = m, kc = k)                      (conv forward / JVP: rows = output pixels)
//   mode 2:  A[m][k] = patch(r = k, kc = m)                      (weight gradient: A = patches^T, contraction over pixels)
//   mode 3:  A[m][k] = D[(mz, y, x), co]  with m = input pixel (mz, yi, xi), k = (dy, dx, co),
//            y * stride = yi + pad_h - dy, x * stride = xi + pad_w - dx   (delta back-propagation = transposed conv; C = cout
//            of the source image D [*, Ho, Wo, C])
// Elements that fall into the zero padding (or between strides) read as 0.
struct ConvGather {
  int mode = 0;
  int Hi = 0, Wi = 0, C = 0, pad_h = 0, pad_w = 0, stride = 1, kh = 1, kw = 1, Ho = 0, Wo = 0;
  FastDiv dC, dkw, dWo, dHo, dWi, dHi;
  void finalize() {
    dC = FastDiv((uint32_t)C); dkw = FastDiv((uint32_t)kw); dWo = FastDiv((uint32_t)Wo); dHo = FastDiv((uint32_t)Ho);
    dWi = FastDiv((uint32_t)Wi); dHi = FastDiv((uint32_t)Hi);
  }
};

struct GemmOperand {
  const float* ptr = nullptr;
  int64_t sz = 0;   // batch stride
  int64_t s0 = 0;   // stride of the row index (m for A, k for B)
  int64_t s1 = 0;   // stride of the col index (k for A, n for B)
  ConvGather conv;  // A operands only: mode != 0 replaces the (s0, s1) addressing by the patch gather above
};

struct GemmEpilogue {
  float scale = 1.f;
  const float* bias = nullptr;  int64_t bias_sz = 0;              // [z][n]
  const float* mask = nullptr;  int64_t mask_sm = 0;              // [m][n], shared by all z
  const float* add = nullptr;   int64_t add_sz = 0; float add_scale = 0.f;  // same m/n strides as C
  int act = -1;                 // -1 none, else lip_op activation
  float* dphi_out = nullptr;    // same layout as C (only with act >= 0)
  float* C_lo = nullptr;        // SIMT path: if set, C receives tf32(v) and C_lo the tf32 remainder (feeds a tcgen05 GEMM)
};

struct GemmProblem {
  int64_t M = 0, N = 0, K = 0, batch = 1;
  GemmOperand A1, B1, A2, B2;   // A2/B2 optional (ptr == nullptr)
  int64_t K2 = 0;
  float* C = nullptr; int64_t c_sz = 0, c_sm = 0;  // n stride is 1
  GemmEpilogue epi;
  // optional split-K scratch (floats): long-K, few-tile problems (per-probe conv weight gradients: K = points x pixels)
  // are cut into K slices whose partial tiles land here and are reduced, in a fixed order, by a second kernel that
  // applies the epilogue.  Deterministic (no atomics).  Ignored when the problem already fills the GPU.
  float* splitk_ws = nullptr; int64_t splitk_ws_elems = 0;
};

int gemm_simt(const GemmProblem& p, cudaStream_t stream);

// tcgen05 3xTF32 path (lip_gemm_tc.cu).  x = hi + lo with hi = tf32_rna(x), lo = tf32_rna(x - hi).
struct TcOperand {
  const float* hi = nullptr;
  const float* lo = nullptr;
  int64_t sz = 0;      // batch stride (elements), multiple of 4
  int64_t ld = 0;      // leading dimension (elements), multiple of 4
  int major_k = 1;     // 1: contraction index contiguous ("K-major"), 0: M/N index contiguous
  // optional device flag written by tf32_split*: 0 = every lo element of this operand is zero (the data is exactly
  // TF32-representable, e.g. +-1 Rademacher probes or one-hot blocks) -> the kernels skip its lo tile loads and the
  // hi x lo MMAs, an exact saving of one third of the tensor work of that operand pair.  Honoured for B1 only.
  const int* lo_nz = nullptr;
};
struct TcGemmProblem {
  int64_t M = 0, N = 0, K = 0, K2 = 0, batch = 1;
  TcOperand A1, B1, A2, B2;   // second pair optional
  int a_batched = 1, b_batched = 1, a2_batched = 1, b2_batched = 1;
  float* C = nullptr; int64_t c_sz = 0, c_sm = 0;
  float* C_lo = nullptr;      // optional: C receives tf32-hi(v), C_lo the remainder (feeds the next GEMM)
  // optional: column sums of the stored values per 32-row block, colsum[z*colsum_sz + (m/32)*colsum_ld + n]
  // (ceil(M/128)*4 slots per z; fuses the bias gradient into the delta-backprop epilogue)
  float* colsum = nullptr; int64_t colsum_sz = 0, colsum_ld = 0;
  GemmEpilogue epi;           // scale, mask OR add (and bias on the single-CTA kernel); no activation
};
bool tc_available();
int gemm_tc(const TcGemmProblem& p, cudaStream_t stream);
// x -> (hi, lo) TF32 split (round-to-nearest hi) into a padded destination: dst[r*ld_dst + c] for r<rows, c<cols
int tf32_split(const float* src, int64_t ld_src, float* hi, float* lo, int64_t ld_dst, int64_t rows,
               int64_t cols, cudaStream_t stream);
// batched form: element (z, r, c) at z*sz + r*ld + c
int tf32_split3(const float* src, int64_t sz_src, int64_t ld_src, float* hi, float* lo, int64_t sz_dst,
                int64_t ld_dst, int64_t batch, int64_t rows, int64_t cols, cudaStream_t stream, int* lo_nz = nullptr);
__device__ __forceinline__ float tf32_round(float x) {
  uint32_t r;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
  return __uint_as_float(r);
}

__device__ __forceinline__ float act_apply(int act, float h, float* dphi) {
  float a, d;
  if (act == LIP_OP_TANH) {
    a = tanhf(h);
    d = 1.f - a * a;
  } else if (act == LIP_OP_RELU) {
    a = h > 0.f ? h : 0.f;
    d = h > 0.f ? 1.f : 0.f;
  } else {  // LIP_OP_GELU_TANH
    const float c = 0.7978845608028654f;
    float u = c * (h + 0.044715f * h * h * h);
    float t = tanhf(u);
    a = 0.5f * h * (1.f + t);
    d = 0.5f * (1.f + t) + 0.5f * h * (1.f - t * t) * c * (1.f + 3.f * 0.044715f * h * h);
  }
  *dphi = d;
  return a;
}

}  // namespace lip
