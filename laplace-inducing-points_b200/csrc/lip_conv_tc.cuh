// lip_conv_tc.cuh — host interface of the tcgen05 implicit-GEMM convolutions (lip_conv_tc.cu).
#pragma once
#include "lip_common.cuh"

namespace lip {

// 1x1 / 3x3 conv with stride 1 or 2 on an [*, H, W, C] image with N output channels: are the three conv GEMMs eligible for the
// tcgen05 path?  (C, N multiples of 32; 128- and 32-pixel runs of the output grid are boxes of whole rows / images;
// LIP_CONV_TC=0 disables the path, LIP_CONV_TC_S2=0 only its strided form)
bool conv_tc_supported(int H, int W, int C, int N, int kh, int kw, int stride, int pad);

// NHWC image batch, pre-split into TF32 (hi, lo): [batch?][imgs][H][W][C]
struct ConvTcImage {
  const float* hi = nullptr;
  const float* lo = nullptr;
  int batched = 0;          // 1: one image stack per batch entry (probe), 0: shared by all batch entries
};

// Fused BatchNorm(eval)-JVP epilogue of a conv unit (lip_resnet.cu header): with acc = the dual-K conv sum and i = (row, n),
//   v = mask[i] * ( g[n] * (acc + pre[z][i]) + xhat[i] * dscale[z][n] + dbeta[z][n] + skip_hi[z][i] + skip_lo[z][i] )
// stored as the TF32 pair (C_out, C_lo) that the next conv's TMA reads directly.  xhat / mask: [rows, N]; skip: layout of C.
struct ConvBnEpilogue {
  int on = 0;
  const float* g = nullptr;
  const float* xhat = nullptr;
  const float* mask = nullptr;                                   // optional (units without relu)
  const float* dscale = nullptr; const float* dbeta = nullptr;   // [z * pstride + n]
  long long pstride = 0;
  const float* skip_hi = nullptr; const float* skip_lo = nullptr;   // optional
  const float* pre = nullptr;      // optional fp32 [layout of C]: added to acc (the probe-folded first JVP term)
};

// out[z][(img, y, x)][n] = sum_{dy,dx,c} A1[z?][img, y + s(dy), x + s(dx), c] * B1[z?][(dy,dx,c)][n]   (+ the same with A2, B2)
//   s(d) = d - pad   (conv forward / JVP)        or        s(d) = pad - d   (transposed = delta back-propagation)
struct ConvTcProblem {
  int64_t imgs = 0, batch = 1;
  int H = 0, W = 0, C = 0, N = 0, kh = 1, kw = 1, pad = 0, transposed = 0;   // H, W: the gathered image
  int stride = 1;                     // 2: rows are the (H/2) x (W/2) output pixels (forward only)
  ConvTcImage A1, A2;                 // A2 optional
  TcOperand B1, B2;                   // [kh*kw*C, N] row-major (MN-major), (hi, lo)
  int b1_batched = 0, b2_batched = 0;
  float* C_out = nullptr; int64_t c_sz = 0, c_sm = 0;
  float* C_lo = nullptr;              // optional (hi, lo) output
  GemmEpilogue epi;                   // scale, bias, mask, add
  ConvBnEpilogue bn;                  // bn.on: replaces `epi` (needs C_lo, c_sm == N)
  // 1: A1 is shared and B1 per probe (no second pair): 128 / N probes share one 128-column tile, so the image tile is
  // staged and read once for all of them (N = 32 / 64; the narrow-N kernels are bound by the A operand's shared-memory reads)
  int fold_probes = 0;
};
int conv_tc(const ConvTcProblem& p, cudaStream_t stream);

// out[z][(dy,dx,c)][n] = scale * sum_{img,y,x} X[img, y + dy - pad, x + dx - pad, c] * D[z][(img,y,x)][n] + add_scale * add[z][..]
struct ConvWgradTcProblem {
  int64_t imgs = 0, batch = 1;
  int H = 0, W = 0, C = 0, N = 0, kh = 1, kw = 1, pad = 0, stride = 1;
  const float* X_hi = nullptr; const float* X_lo = nullptr;   // shared image [imgs, H, W, C]
  TcOperand D;                                                // deltas [batch][imgs*H*W, N] (hi, lo), MN-major
  float* C_out = nullptr; int64_t c_sz = 0, c_sm = 0;
  GemmEpilogue epi;                                           // scale, add
  float* ws = nullptr; int64_t ws_elems = 0;                  // split-K scratch: conv_wgrad_tc_splits(..) * batch * kh*kw*C * N floats
};
int64_t conv_wgrad_tc_splits(int64_t imgs, int Ho, int Wo, int C, int N, int kh, int kw, int64_t batch);
// zero-upsampled copy of a (hi, lo) delta image [images, Ho, Wo, C] -> [images, 2Ho, 2Wo, C]: the delta back-propagation of a
// stride-2 conv is the stride-1 transposed conv of this image
int conv_tc_upsample2(const float* d_hi, const float* d_lo, float* up_hi, float* up_lo, int64_t images, int Ho, int Wo, int C,
                      cudaStream_t stream);
int conv_wgrad_tc(const ConvWgradTcProblem& p, cudaStream_t stream);

}  // namespace lip
