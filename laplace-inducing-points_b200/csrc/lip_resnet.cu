// lip_resnet.cu — probe-batched GGN / W / W^T operators for residual conv programs (ResNet1M, src/scalemodels.py:70-157).
//
// Same operator semantics as lip_model.cu (src/ggn.py:9-146).  The network is a chain of conv + BatchNorm(eval) units
//   y = relu?( scale * xhat + beta + skip ),   xhat = (conv(x, W) - mean) * rsqrt(var + eps),   g = scale * rsqrt(var + eps)
// (stem; per BasicBlock: conv-bn-relu, conv-bn, optional 1x1 strided conv-bn on the residual, add, relu), a global mean and
// a Dense head.  BatchNorm scale / bias ARE parameters (they are in the flat vector); mean / var are constants (ggn.py:52).
//   JVP  unit:  dY[b] = mask * ( g * (patches(X) . dW[b] + patches(T_src[b]) . W) + xhat * dscale[b] + dbeta[b] + T_skip[b] )
//   VJP  unit:  Dy = Dout * mask;  gbeta = colsum Dy;  gscale = colsum (Dy * xhat);  cot[skip] (+)= Dy;  Dh = g * Dy;
//               gW[b] = patches(X)^T . Dh[b];  cot[src] (+)= transposed_conv(Dh[b], W)
// Convs run as IMPLICIT GEMMs: the A operand gathers its im2col patches straight from the NHWC image - the cached input
// activation, the probe's tangent image, or (transposed conv) the delta - so no patch buffer, im2col or col2im pass exists and
// memory stays O(activations).  Stride-1 convs with >= 32 channels run on the tensor cores (tcgen05 3xTF32, TMA box loads
// from the image, lip_conv_tc.cu); the 3-channel stem and the strided convs on the fp32 SIMT kernels (ConvGather in
// lip_common.cuh).
// Tangents / cotangents live in four rotating [B, M, H, W, C] slots (block input, branch, shortcut, block output).
#include <stdlib.h>

#include <algorithm>

#include <new>
#include <vector>

#include "lip_model.cuh"
#include "lip_conv_tc.cuh"

using namespace lip;

void lip_model::free_resnet_cache() {
  for (auto& u : RB) {
    if (u.Xin) cudaFree(u.Xin);
    if (u.Wt) cudaFree(u.Wt);
    if (u.xhat) cudaFree(u.xhat);
    if (u.mask) cudaFree(u.mask);
    if (u.g) cudaFree(u.g);
    for (float* q : {u.Xh, u.Xl, u.Wh, u.Wl, u.Wth, u.Wtl}) if (q) cudaFree(q);
    u.Xin = u.Wt = u.xhat = u.mask = u.g = nullptr;
    u.Xh = u.Xl = u.Wh = u.Wl = u.Wth = u.Wtl = nullptr;
    u.tc = false;
  }
  if (rn_mean_act) cudaFree(rn_mean_act);
  rn_mean_act = nullptr;
}

namespace {

constexpr float BN_EPS = 1e-5f;   // flax.linen.BatchNorm default epsilon

inline unsigned ew_grid(long long total) {
  long long g = (total + 255) / 256;
  if (g > 148ll * 32) g = 148ll * 32;
  if (g < 1) g = 1;
  return (unsigned)g;
}

// bind time: h = conv output [R, C] -> xhat, g;  y = relu?(scale * xhat + beta + skip) -> out, mask
__global__ void bn_fwd_kernel(const float* __restrict__ h, const float* __restrict__ mean, const float* __restrict__ var,
                              const float* __restrict__ scale, const float* __restrict__ beta, const float* __restrict__ skip,
                              float* __restrict__ xhat, float* __restrict__ mask, float* __restrict__ g, float* __restrict__ out,
                              long long total, int C, int relu) {
  for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total; idx += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(idx % C);
    const float rstd = rsqrtf(var[c] + BN_EPS);
    const float xh = (h[idx] - mean[c]) * rstd;
    xhat[idx] = xh;
    if (idx < C) g[c] = scale[c] * rstd;
    float y = fmaf(scale[c], xh, beta[c]);
    if (skip) y += skip[idx];
    if (relu) {
      mask[idx] = y > 0.f ? 1.f : 0.f;
      y = fmaxf(y, 0.f);
    }
    out[idx] = y;
  }
}

// dY[b][i] = mask[i] * ( g[c] * dH[b][i] + xhat[i] * dscale[b][c] + dbeta[b][c] + Tskip[b][i] ),  i in [0, per_z)
// out_lo != null: the result is stored as a TF32 (hi, lo) pair (and Tskip read as one) for the tcgen05 conv GEMMs
__global__ void bn_jvp_kernel(const float* __restrict__ dH, const float* __restrict__ g, const float* __restrict__ xhat,
                              const float* __restrict__ mask, const float* __restrict__ dscale, const float* __restrict__ dbeta,
                              long long pstride, const float* __restrict__ tskip, const float* __restrict__ tskip_lo,
                              float* __restrict__ out, float* __restrict__ out_lo, long long total, long long per_z, int C) {
  for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total; idx += (long long)gridDim.x * blockDim.x) {
    const long long z = idx / per_z, i = idx % per_z;
    const int c = (int)(i % C);
    float v = fmaf(g[c], dH[idx], fmaf(__ldg(xhat + i), __ldg(dscale + z * pstride + c), __ldg(dbeta + z * pstride + c)));
    if (tskip) v += tskip_lo ? tskip[idx] + tskip_lo[idx] : tskip[idx];
    if (mask) v *= __ldg(mask + i);
    if (out_lo) {
      const float hh = tf32_round(v);
      out[idx] = hh;
      out_lo[idx] = tf32_round(v - hh);
    } else {
      out[idx] = v;
    }
  }
}

// BatchNorm backward of one unit in ONE pass over Dout (deterministic, no atomics).  Grid (chunks, probes), blockDim.x a multiple
// of C so that every thread sees one column of the flat [R * C] arrays; each CTA owns a slice of rows:
//   Dy = Dout * mask;   cot_skip (+)= Dy;   Dh = g[c] * Dy  (fp32, or a TF32 (hi, lo) pair for the tcgen05 conv GEMMs)
//   part[b][chunk][0][c] = sum_r Dy,   part[b][chunk][1][c] = sum_r Dy * xhat       (stage 1 of the parameter gradients)
__global__ void bn_backward_kernel(const float* __restrict__ dout, const float* __restrict__ mask, const float* __restrict__ xhat,
                                   const float* __restrict__ g, float* __restrict__ dh, float* __restrict__ dh_lo, float* cot_skip,
                                   long long per_z, int C, float* __restrict__ part) {
  extern __shared__ float sm[];
  float* s1 = sm;
  float* s2 = sm + blockDim.x;
  const long long b = blockIdx.y;
  const int nch = gridDim.x;
  const float* d = dout + b * per_z;
  float* oh = dh + b * per_z;
  float* ol = dh_lo ? dh_lo + b * per_z : nullptr;
  float* os = cot_skip ? cot_skip + b * per_z : nullptr;
  // slices are multiples of blockDim.x (itself a multiple of C): the thread <-> column mapping is the same in every slice
  const long long per_chunk = ((per_z + nch - 1) / nch + blockDim.x - 1) / blockDim.x * blockDim.x;
  const long long lo = blockIdx.x * per_chunk, hi = lo + per_chunk < per_z ? lo + per_chunk : per_z;
  const float gc = g[threadIdx.x % C];
  float a1 = 0.f, a2 = 0.f;
  constexpr int U = 4;
  long long i = lo + threadIdx.x;
  for (; i + (U - 1) * (long long)blockDim.x < hi; i += U * (long long)blockDim.x) {
    float dy[U], xh[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const long long j = i + u * (long long)blockDim.x;
      dy[u] = __ldg(d + j);
      if (mask) dy[u] *= __ldg(mask + j);
      xh[u] = __ldg(xhat + j);
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const long long j = i + u * (long long)blockDim.x;
      a1 += dy[u];
      a2 = fmaf(dy[u], xh[u], a2);
      if (os) os[j] = dy[u];
      const float v = gc * dy[u];
      if (ol) {
        const float hh = tf32_round(v);
        oh[j] = hh;
        ol[j] = tf32_round(v - hh);
      } else {
        oh[j] = v;
      }
    }
  }
  for (; i < hi; i += blockDim.x) {
    float dy = __ldg(d + i);
    if (mask) dy *= __ldg(mask + i);
    a1 += dy;
    a2 = fmaf(dy, __ldg(xhat + i), a2);
    if (os) os[i] = dy;
    const float v = gc * dy;
    if (ol) {
      const float hh = tf32_round(v);
      oh[i] = hh;
      ol[i] = tf32_round(v - hh);
    } else {
      oh[i] = v;
    }
  }
  s1[threadIdx.x] = a1; s2[threadIdx.x] = a2;
  __syncthreads();
  if ((int)threadIdx.x < C) {
    float t1 = 0.f, t2 = 0.f;
    for (int k = threadIdx.x; k < (int)blockDim.x; k += C) { t1 += s1[k]; t2 += s2[k]; }
    float* p = part + ((b * nch + blockIdx.x) * 2) * C;
    p[threadIdx.x] = t1;
    p[C + threadIdx.x] = t2;
  }
}
// Stage 2: gbeta[b][c] / gscale[b][c] = scale * sum_chunk part + add_scale * add
__global__ void bn_param_grad_reduce_kernel(const float* __restrict__ part, int nch, int C, float* __restrict__ out_beta,
                                            float* __restrict__ out_scale, long long out_sz, float scale,
                                            const float* __restrict__ add_beta, const float* __restrict__ add_scale_p,
                                            long long add_sz, float add_scale) {
  const long long b = blockIdx.x;
  const int c = threadIdx.x;
  if (c >= C) return;
  float t1 = 0.f, t2 = 0.f;
  for (int k = 0; k < nch; ++k) {
    const float* p = part + ((b * nch + k) * 2) * C;
    t1 += p[c]; t2 += p[C + c];
  }
  float vb = scale * t1, vs = scale * t2;
  if (add_beta) { vb += add_scale * add_beta[b * add_sz + c]; vs += add_scale * add_scale_p[b * add_sz + c]; }
  out_beta[b * out_sz + c] = vb;
  out_scale[b * out_sz + c] = vs;
}

// Wt[(tap*cout + co)*cin + ci] = W[(tap*cin + ci)*cout + co]
__global__ void conv_wt_kernel(const float* __restrict__ W, float* __restrict__ Wt, int taps, int cin, int cout) {
  const int total = taps * cin * cout;
  for (int idx = blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += gridDim.x * blockDim.x) {
    const int ci = idx % cin, t = idx / cin, co = t % cout, tap = t / cout;
    Wt[idx] = W[(tap * cin + ci) * cout + co];
  }
}

// out[mz][c] = mean over hw of in[mz][hw][c]
__global__ void global_mean_kernel(const float* __restrict__ in, const float* __restrict__ in_lo, float* __restrict__ out,
                                   long long total, int HW, int C) {
  for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total; idx += (long long)gridDim.x * blockDim.x) {
    const long long mz = idx / C;
    const int c = (int)(idx % C);
    const float* p = in + mz * HW * C + c;
    float acc = 0.f;
    for (int i = 0; i < HW; ++i) acc += p[(long long)i * C];
    if (in_lo) {                    // TF32 (hi, lo) pair: add the lo parts
      const float* q = in_lo + mz * HW * C + c;
      float acc2 = 0.f;
      for (int i = 0; i < HW; ++i) acc2 += q[(long long)i * C];
      acc += acc2;
    }
    out[idx] = acc / (float)HW;
  }
}
// cot_in[mz][hw][c] = cot_out[mz][c] / HW
__global__ void global_mean_bwd_kernel(const float* __restrict__ g, float* __restrict__ out, long long total, int HW, int C) {
  const float inv = 1.f / (float)HW;
  for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total; idx += (long long)gridDim.x * blockDim.x) {
    const long long mz = idx / ((long long)HW * C);
    const int c = (int)(idx % C);
    out[idx] = inv * __ldg(g + mz * C + c);
  }
}

// ---- gradients with respect to the input images Z (lip_zgrad for residual programs) ---------------------------------
// One unit's step of the reverse pass (see resnet_zgrad):  ebar = mask * E,  qbar = mask * Q;  skip cotangents stored;
//   rawE = g * ebar;   rawQ = g * qbar + rstd * dscale[b] * ebar      (the BatchNorm-scale tangent x xhat(Z) term)
__global__ void rn_zgrad_bn_kernel(const float* __restrict__ E, const float* __restrict__ Q, const float* __restrict__ mask,
                                   const float* __restrict__ g, const float* __restrict__ var, const float* __restrict__ dscale,
                                   long long pstride, float* __restrict__ eskip, float* __restrict__ qskip, float* __restrict__ rawE,
                                   float* __restrict__ rawQ, long long total, long long per_z, int C) {
  for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total; idx += (long long)gridDim.x * blockDim.x) {
    const long long z = idx / per_z, i = idx % per_z;
    const int c = (int)(i % C);
    float e = E[idx], q = Q[idx];
    if (mask) { const float mk = __ldg(mask + i); e *= mk; q *= mk; }
    if (eskip) { eskip[idx] = e; qskip[idx] = q; }
    const float gc = __ldg(g + c);
    rawE[idx] = gc * e;
    rawQ[idx] = fmaf(gc, q, rsqrtf(__ldg(var + c) + BN_EPS) * __ldg(dscale + z * pstride + c) * e);
  }
}
// per-probe re-lay of the tangent kernels for the transposed conv: Wt[b][(tap*cout + co)*cin + ci] = V[b][woff + (tap*cin + ci)*cout + co]
__global__ void conv_wt_batched_kernel(const float* __restrict__ V, long long vstride, float* __restrict__ Wt, int taps, int cin, int cout) {
  const int total = taps * cin * cout;
  const float* W = V + (long long)blockIdx.y * vstride;
  float* o = Wt + (long long)blockIdx.y * total;
  for (int idx = blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += gridDim.x * blockDim.x) {
    const int ci = idx % cin, t = idx / cin, co = t % cout, tap = t / cout;
    o[idx] = W[(tap * cin + ci) * cout + co];
  }
}

struct RnWs {
  float* slot[4];   // tangent / cotangent tensors [B, M, slot_elems]
  float* slot_lo[4];// pairs mode (lip_model::rn_pairs): TF32 lo parts of the JVP tangents (slot[] then holds the hi parts)
  float* raw;       // conv GEMM output dH / Dh, [B, max R*cout]
  float* raw_lo;    // tensor path: TF32 lo part of Dh (raw then holds the hi part)
  float* img_h;     // tensor path: TF32 (hi, lo) split of the tangent image a conv reads, [B, M, slot_elems] each
  float* img_l;
  float* vh;        // tensor path: TF32 (hi, lo) split of one unit's tangent kernels, [B, max Kc*cout] each
  float* vl;
  float* col;       // split-K scratch of the per-probe kernel-gradient GEMMs
  float* head;      // [B, M, C_last] mean tangent / cotangent  +  [B, M, K] delta at the logits
  float* dl;
};
struct RnSizes { size_t slot, raw, col, head, dl, raw_lo, img, vsplit, slot_lo; };

RnSizes rn_sizes(const lip_model* m, int64_t B) {
  RnSizes z{};
  z.slot = align_up((size_t)B * m->M * (size_t)m->rn_slot_elems, 64);
  size_t raw = 0, col = 0, vs = 0, img = (size_t)B * m->M * (size_t)m->rn_slot_elems;
  bool any_tc = false;
  for (const ConvBN& u : m->RB) {
    const size_t r = (size_t)B * m->M * u.P() * u.cout;
    raw = r > raw ? r : raw;
    size_t slices = 8;                                       // up to 8 K-slices of every probe's [Kc x cout] gradient (SIMT)
    if (u.tc) {
      any_tc = true;
      const size_t s = (size_t)conv_wgrad_tc_splits(m->M, u.Ho, u.Wo, u.cin, u.cout, u.kh, u.kw, B);
      slices = s > slices ? s : slices;
      const size_t v = (size_t)B * u.Kc() * u.cout;
      vs = v > vs ? v : vs;
      // a strided unit back-propagates through the zero-upsampled delta image [B, M, Hi, Wi, cout]
      if (u.stride == 2) { const size_t e = (size_t)B * m->M * u.Hi * u.Wi * u.cout; img = e > img ? e : img; }
    }
    const size_t c = (size_t)B * u.Kc() * u.cout * slices;
    col = c > col ? c : col;
  }
  z.raw = align_up(raw, 64); z.col = align_up(col, 64);
  z.raw_lo = any_tc ? z.raw : 0;
  z.img = any_tc ? align_up(img, 64) : 0;
  z.vsplit = align_up(vs, 64);
  z.slot_lo = m->rn_pairs ? z.slot : 0;
  z.head = align_up((size_t)B * m->M * m->rn_C, 64);
  z.dl = align_up((size_t)B * m->M * m->K, 64);
  return z;
}

int rn_carve(const lip_model* m, int64_t B, void* ws, size_t bytes, RnWs* w) {
  const size_t need = resnet_ws_bytes(m, B);
  if (bytes < need || ws == nullptr) {
    set_error("workspace too small: need %zu bytes, got %zu", need, bytes);
    return LIP_ERR_WORKSPACE;
  }
  const RnSizes z = rn_sizes(m, B);
  float* p = (float*)align_up((uintptr_t)ws, 256);
  for (int i = 0; i < 4; ++i) { w->slot[i] = p; p += z.slot; }
  w->raw = p; p += z.raw;
  w->col = p; p += z.col;
  w->head = p; p += z.head;
  w->dl = p; p += z.dl;
  w->raw_lo = p; p += z.raw_lo;
  w->img_h = p; p += z.img;
  w->img_l = p; p += z.img;
  w->vh = p; p += z.vsplit;
  w->vl = p; p += z.vsplit;
  for (int i = 0; i < 4; ++i) { w->slot_lo[i] = p; p += z.slot_lo; }
  return LIP_OK;
}

// ---- JVP sweep: V[B, D] -> dlogits [B, M, K] in dst --------------------------------------------------------------------
int rn_jvp_sweep(lip_model* m, const float* V, int64_t B, const RnWs& w, float* dst, cudaStream_t st) {
  if (m->lo_nz) LIP_CHECK_CUDA(cudaMemsetAsync(m->lo_nz, 0, sizeof(int) * m->RB.size(), st));
  const bool pairs = m->rn_pairs;
  for (const ConvBN& u : m->RB) {
    const int64_t R = m->M * u.P(), Kc = u.Kc();
    int rc;
    if (u.tc) {
      // tensor path: split the probe's tangent kernels (and, outside pairs mode, the tangent image) into TF32 (hi, lo), then
      // one dual-K implicit GEMM; in pairs mode its epilogue is the whole BatchNorm-JVP and writes the next (hi, lo) tangent
      const int64_t img_elems = m->M * (int64_t)u.Hi * u.Wi * u.cin;
      int* nz = m->lo_nz + (&u - m->RB.data());
      rc = tf32_split3(V + u.woff, m->D, Kc * u.cout, w.vh, w.vl, Kc * u.cout, Kc * u.cout, B, 1, Kc * u.cout, st, nz);
      if (rc) return rc;
      if (!pairs) {
        rc = tf32_split3(w.slot[u.src], 0, B * img_elems, w.img_h, w.img_l, 0, B * img_elems, 1, 1, B * img_elems, st);
        if (rc) return rc;
      }
      ConvTcProblem c;
      c.imgs = m->M; c.batch = B; c.H = u.Hi; c.W = u.Wi; c.C = u.cin; c.N = u.cout; c.kh = u.kh; c.kw = u.kw; c.pad = u.pad_h;
      c.stride = u.stride;
      c.A1.hi = u.Xh; c.A1.lo = u.Xl; c.A1.batched = 0;
      c.B1.hi = w.vh; c.B1.lo = w.vl; c.B1.sz = Kc * u.cout; c.B1.ld = u.cout; c.B1.major_k = 0; c.B1.lo_nz = nz; c.b1_batched = 1;
      c.A2.hi = pairs ? w.slot[u.src] : w.img_h; c.A2.lo = pairs ? w.slot_lo[u.src] : w.img_l; c.A2.batched = 1;
      c.B2.hi = u.Wh; c.B2.lo = u.Wl; c.B2.sz = 0; c.B2.ld = u.cout; c.B2.major_k = 0; c.b2_batched = 0;
      c.c_sz = R * (int64_t)u.cout; c.c_sm = u.cout;
      if (pairs) {
        if (u.cout <= 64 && B >= 2) {
          // narrow layers are bound by the image tile's shared-memory reads: the first term (shared image x per-probe kernels)
          // runs with 128 / cout probes folded into one tile -> raw, the second term adds it in its BatchNorm epilogue
          ConvTcProblem f = c;
          f.A2 = ConvTcImage(); f.B2 = TcOperand();
          f.fold_probes = 1;
          f.C_out = w.raw;
          rc = conv_tc(f, st);
          if (rc) return rc;
          c.A1 = c.A2; c.B1 = c.B2; c.b1_batched = 0;
          c.A2 = ConvTcImage(); c.B2 = TcOperand();
          c.bn.pre = w.raw;
        }
        c.C_out = w.slot[u.dst]; c.C_lo = w.slot_lo[u.dst];
        c.bn.on = 1; c.bn.g = u.g; c.bn.xhat = u.xhat; c.bn.mask = u.mask;
        c.bn.dscale = V + u.scale_off; c.bn.dbeta = V + u.beta_off; c.bn.pstride = m->D;
        if (u.skip >= 0) { c.bn.skip_hi = w.slot[u.skip]; c.bn.skip_lo = w.slot_lo[u.skip]; }
        rc = conv_tc(c, st);
        if (rc) return rc;
        continue;
      }
      c.C_out = w.raw;
      rc = conv_tc(c, st);
    } else if (resnet_stem_fusable(u)) {
      rc = resnet_stem_jvp(m, u, V, m->D, w.slot[u.dst], pairs ? w.slot_lo[u.dst] : nullptr, B, st);
      if (rc) return rc;
      continue;
    } else {
      GemmProblem p;
      p.M = R; p.N = u.cout; p.K = Kc; p.batch = B;
      p.A1.ptr = u.Xin; p.A1.sz = 0; p.A1.conv = u.gather(1);                    // patches of the cached activation
      p.B1 = {V + u.woff, m->D, u.cout, 1};
      if (u.src != -2) {
        p.A2.ptr = w.slot[u.src]; p.A2.sz = m->M * (int64_t)u.Hi * u.Wi * u.cin;  // patches of the probe's tangent image
        p.A2.conv = u.gather(1);
        p.B2 = {m->theta + u.woff, 0, u.cout, 1};
        p.K2 = Kc;
      }
      p.C = w.raw; p.c_sz = R * (int64_t)u.cout; p.c_sm = u.cout;
      rc = gemm_simt(p, st);
    }
    if (rc) return rc;
    const long long per_z = R * (long long)u.cout, total = per_z * B;
    bn_jvp_kernel<<<ew_grid(total), 256, 0, st>>>(w.raw, u.g, u.xhat, u.mask, V + u.scale_off, V + u.beta_off, m->D,
                                                  u.skip >= 0 ? w.slot[u.skip] : nullptr,
                                                  (pairs && u.skip >= 0) ? w.slot_lo[u.skip] : nullptr, w.slot[u.dst],
                                                  pairs ? w.slot_lo[u.dst] : nullptr, total, per_z, u.cout);
    LIP_LAUNCH_CHECK();
  }
  const int HW = m->rn_H * m->rn_W, C = m->rn_C;
  {
    const long long total = (long long)B * m->M * C;
    global_mean_kernel<<<ew_grid(total), 256, 0, st>>>(w.slot[m->rn_last_slot], pairs ? w.slot_lo[m->rn_last_slot] : nullptr, w.head,
                                                       total, HW, C);
    LIP_LAUNCH_CHECK();
  }
  GemmProblem p;   // head: dlogits[b] = mean_act . dWd[b] + Tmean[b] . Wd + dbd[b]
  p.M = m->M; p.N = m->K; p.K = C; p.batch = B;
  p.A1 = {m->rn_mean_act, 0, C, 1};
  p.B1 = {V + m->rn_dense_woff, m->D, m->K, 1};
  p.A2 = {w.head, m->M * (int64_t)C, C, 1};
  p.B2 = {m->theta + m->rn_dense_woff, 0, m->K, 1};
  p.K2 = C;
  p.C = dst; p.c_sz = m->M * (int64_t)m->K; p.c_sm = m->K;
  p.epi.bias = V + m->rn_dense_boff; p.epi.bias_sz = m->D;
  return gemm_simt(p, st);
}

// ---- VJP sweep: dl [B, M, K] -> out[B, D] = scale * J^T dl + add_scale * add ---------------------------------------------
int rn_vjp_sweep(lip_model* m, const float* dl, int64_t B, const RnWs& w, float* out, float scale, const float* add,
                 float add_scale, cudaStream_t st) {
  const int HW = m->rn_H * m->rn_W, C = m->rn_C;
  const size_t col_elems = rn_sizes(m, B).col;
  {  // head
    GemmProblem p;
    p.M = C; p.N = m->K; p.K = m->M; p.batch = B;
    p.A1 = {m->rn_mean_act, 0, 1, C};
    p.B1 = {dl, m->M * (int64_t)m->K, m->K, 1};
    p.C = out + m->rn_dense_woff; p.c_sz = m->D; p.c_sm = m->K;
    p.epi.scale = scale;
    if (add) { p.epi.add = add + m->rn_dense_woff; p.epi.add_sz = m->D; p.epi.add_scale = add_scale; }
    int rc = gemm_simt(p, st);
    if (rc) return rc;
    rc = launch_bias_grad(dl, nullptr, m->M, m->K, m->K, B, out + m->rn_dense_boff, m->D, scale,
                          add ? add + m->rn_dense_boff : nullptr, m->D, add_scale, st);
    if (rc) return rc;
    GemmProblem q;   // cotangent of the mean activations: [M x C] = dl [M x K] . Wd^T
    q.M = m->M; q.N = C; q.K = m->K; q.batch = B;
    q.A1 = {dl, m->M * (int64_t)m->K, m->K, 1};
    q.B1 = {m->theta + m->rn_dense_woff, 0, 1, m->K};
    q.C = w.head; q.c_sz = m->M * (int64_t)C; q.c_sm = C;
    rc = gemm_simt(q, st);
    if (rc) return rc;
    const long long total = (long long)B * m->M * HW * C;
    global_mean_bwd_kernel<<<ew_grid(total), 256, 0, st>>>(w.head, w.slot[m->rn_last_slot], total, HW, C);
    LIP_LAUNCH_CHECK();
  }
  for (int i = (int)m->RB.size() - 1; i >= 0; --i) {
    const ConvBN& u = m->RB[i];
    const int64_t R = m->M * u.P(), Kc = u.Kc();
    const long long per_z = R * (long long)u.cout, total = per_z * B;
    const float* dout = w.slot[u.dst];
    {  // BatchNorm backward in one pass: Dy -> skip cotangent (first contribution to that slot: plain store), Dh = g * Dy, and
       // the parameter-gradient partials (they land in the split-K scratch, which the kernel-gradient GEMM reuses afterwards)
      const int threads = u.cout * (512 / u.cout > 0 ? 512 / u.cout : 1);
      int nch = (int)(per_z / (threads * 16));
      if (nch > 128) nch = 128;
      const long long cap = (long long)(col_elems / ((size_t)B * 2 * u.cout));
      if (nch > cap) nch = (int)cap;
      if (nch < 1) nch = 1;
      dim3 grid((unsigned)nch, (unsigned)B);
      bn_backward_kernel<<<grid, threads, 2 * threads * sizeof(float), st>>>(dout, u.mask, u.xhat, u.g, w.raw,
                                                                            u.tc ? w.raw_lo : nullptr,
                                                                            u.skip >= 0 ? w.slot[u.skip] : nullptr, per_z, u.cout,
                                                                            w.col);
      LIP_LAUNCH_CHECK();
      bn_param_grad_reduce_kernel<<<(unsigned)B, u.cout <= 32 ? 32 : (u.cout + 31) / 32 * 32, 0, st>>>(
          w.col, nch, u.cout, out + u.beta_off, out + u.scale_off, m->D, scale, add ? add + u.beta_off : nullptr,
          add ? add + u.scale_off : nullptr, m->D, add_scale);
      LIP_LAUNCH_CHECK();
    }
    if (u.tc) {
      ConvWgradTcProblem c;   // kernel gradient [Kc x cout] = patches(X)^T . Dh on the tensor cores (split-K)
      c.imgs = m->M; c.batch = B; c.H = u.Hi; c.W = u.Wi; c.C = u.cin; c.N = u.cout; c.kh = u.kh; c.kw = u.kw; c.pad = u.pad_h;
      c.stride = u.stride;
      c.X_hi = u.Xh; c.X_lo = u.Xl;
      c.D.hi = w.raw; c.D.lo = w.raw_lo; c.D.sz = per_z; c.D.ld = u.cout; c.D.major_k = 0;
      c.C_out = out + u.woff; c.c_sz = m->D; c.c_sm = u.cout;
      c.epi.scale = scale;
      if (add) { c.epi.add = add + u.woff; c.epi.add_sz = m->D; c.epi.add_scale = add_scale; }
      c.ws = w.col; c.ws_elems = (int64_t)col_elems;
      int rc = conv_wgrad_tc(c, st);
      if (rc) return rc;
      ConvTcProblem q;        // cotangent of the input image = transposed conv of Dh (zero-upsampled first for a strided unit)
      q.imgs = m->M; q.batch = B; q.H = u.Hi; q.W = u.Wi; q.C = u.cout; q.N = u.cin; q.kh = u.kh; q.kw = u.kw; q.pad = u.pad_h;
      q.transposed = 1;
      q.A1.hi = w.raw; q.A1.lo = w.raw_lo; q.A1.batched = 1;
      if (u.stride == 2) {
        rc = conv_tc_upsample2(w.raw, w.raw_lo, w.img_h, w.img_l, B * m->M, u.Ho, u.Wo, u.cout, st);
        if (rc) return rc;
        q.A1.hi = w.img_h; q.A1.lo = w.img_l;
      }
      q.B1.hi = u.Wth; q.B1.lo = u.Wtl; q.B1.sz = 0; q.B1.ld = u.cin; q.B1.major_k = 0; q.b1_batched = 0;
      q.C_out = w.slot[u.src]; q.c_sz = m->M * (int64_t)u.Hi * u.Wi * u.cin; q.c_sm = u.cin;
      if (u.accumulate) { q.epi.add = q.C_out; q.epi.add_sz = q.c_sz; q.epi.add_scale = 1.f; }
      rc = conv_tc(q, st);
      if (rc) return rc;
      continue;
    }
    {  // kernel gradient [Kc x cout] = patches(X)^T . Dh
      GemmProblem p;
      p.M = Kc; p.N = u.cout; p.K = R; p.batch = B;
      p.A1.ptr = u.Xin; p.A1.sz = 0; p.A1.conv = u.gather(2);
      p.B1 = {w.raw, per_z, u.cout, 1};
      p.C = out + u.woff; p.c_sz = m->D; p.c_sm = u.cout;
      p.epi.scale = scale;
      if (add) { p.epi.add = add + u.woff; p.epi.add_sz = m->D; p.epi.add_scale = add_scale; }
      p.splitk_ws = w.col; p.splitk_ws_elems = (int64_t)col_elems;
      int rc = gemm_simt(p, st);
      if (rc) return rc;
    }
    if (u.src == -2) continue;
    // cotangent of the input image = transposed conv of Dh: one GEMM whose A operand gathers Dh (no col buffer / col2im)
    GemmProblem q;
    q.M = m->M * (int64_t)u.Hi * u.Wi; q.N = u.cin; q.K = (int64_t)u.kh * u.kw * u.cout; q.batch = B;
    q.A1.ptr = w.raw; q.A1.sz = per_z; q.A1.conv = u.gather(3);
    q.B1 = {u.Wt, 0, u.cin, 1};
    q.C = w.slot[u.src]; q.c_sz = q.M * (int64_t)u.cin; q.c_sm = u.cin;
    if (u.accumulate) { q.epi.add = q.C; q.epi.add_sz = q.c_sz; q.epi.add_scale = 1.f; }   // second branch into this slot
    int rc = gemm_simt(q, st);
    if (rc) return rc;
  }
  return LIP_OK;
}

}  // namespace

namespace lip {

int resnet_parse(lip_model* m, const lip_layer_desc* L, int32_t n, int64_t num_params) {
  m->is_resnet = true;
  const lip_layer_desc& in = L[0];
  LIP_REQUIRE(in.kh > 0 && in.kw > 0 && in.in_features > 0, "lip_model_create: INPUT needs height, width, channels > 0");
  m->in_h = in.kh; m->in_w = in.kw; m->in_c = in.in_features;
  int H = in.kh, W = in.kw, C = in.in_features;
  int64_t counted = 0, stats = 0;
  int64_t max_elems = (int64_t)H * W * C;
  int i = 1;
  // conv (+ the BATCHNORM that must follow it) -> one ConvBN unit; (Hs, Ws, Cs) is the unit's input tensor
  auto conv_bn = [&](int ci, int bi, int Hs, int Ws, int Cs, ConvBN* u) -> int {
    const lip_layer_desc& c = L[ci];
    const lip_layer_desc& b = L[bi];
    LIP_REQUIRE(c.in_features == Cs && c.out_features > 0 && c.kh > 0 && c.kw > 0 && (c.stride == 1 || c.stride == 2) && c.pad >= 0,
                "lip_model_create: conv at %d: bad shape (cin %d vs %d, stride %d)", ci, c.in_features, Cs, c.stride);
    LIP_REQUIRE(b.in_features == c.out_features, "lip_model_create: BATCHNORM at %d has %d channels, conv has %d", bi,
                b.in_features, c.out_features);
    u->Hi = Hs; u->Wi = Ws; u->cin = Cs; u->kh = c.kh; u->kw = c.kw; u->stride = c.stride; u->pad_h = u->pad_w = c.pad;
    u->Ho = (Hs + c.stride - 1) / c.stride; u->Wo = (Ws + c.stride - 1) / c.stride;      // SAME
    u->cout = c.out_features;
    u->woff = c.kernel_offset; u->scale_off = b.kernel_offset; u->beta_off = b.bias_offset;
    const int64_t nk = (int64_t)c.kh * c.kw * Cs * c.out_features;
    LIP_REQUIRE(c.kernel_offset >= 0 && c.kernel_offset + nk <= num_params && b.kernel_offset >= 0 && b.bias_offset >= 0 &&
                    b.kernel_offset + c.out_features <= num_params && b.bias_offset + c.out_features <= num_params,
                "lip_model_create: conv/BATCHNORM at %d/%d has an invalid offset", ci, bi);
    LIP_REQUIRE(c.out_features <= 512, "lip_model_create: conv at %d: more than 512 channels", ci);
    u->stats_off = stats;
    stats += 2 * c.out_features;
    counted += nk + 2 * c.out_features;
    const int64_t e = (int64_t)u->Ho * u->Wo * u->cout;
    max_elems = e > max_elems ? e : max_elems;
    return LIP_OK;
  };
  auto is = [&](int k, int op) { return k < n && L[k].op == op; };
  LIP_REQUIRE(is(1, LIP_OP_CONV2D) && is(2, LIP_OP_BATCHNORM) && is(3, LIP_OP_RELU),
              "lip_model_create: a residual program starts INPUT, CONV2D, BATCHNORM, RELU");
  int x = 0;   // slot of the current tensor
  {
    ConvBN u;
    int rc = conv_bn(1, 2, H, W, C, &u);
    if (rc) return rc;
    u.relu = 1; u.src = -2; u.dst = x; u.skip = -1;
    m->RB.push_back(u);
    H = u.Ho; W = u.Wo; C = u.cout;
  }
  i = 4;
  while (is(i, LIP_OP_RES_SAVE)) {
    LIP_REQUIRE(is(i + 1, LIP_OP_CONV2D) && is(i + 2, LIP_OP_BATCHNORM) && is(i + 3, LIP_OP_RELU) && is(i + 4, LIP_OP_CONV2D) &&
                    is(i + 5, LIP_OP_BATCHNORM),
                "lip_model_create: block at %d must be RES_SAVE, CONV2D, BATCHNORM, RELU, CONV2D, BATCHNORM, ...", i);
    int free_slots[3], nf = 0;
    for (int s = 0; s < 4; ++s) if (s != x) free_slots[nf++] = s;
    const int y = free_slots[0], r = free_slots[1], o = free_slots[2];
    ConvBN c0, c1;
    int rc = conv_bn(i + 1, i + 2, H, W, C, &c0);
    if (rc) return rc;
    c0.relu = 1; c0.src = x; c0.dst = y; c0.skip = -1; c0.accumulate = 1;   // x's cotangent already holds the skip branch
    rc = conv_bn(i + 4, i + 5, c0.Ho, c0.Wo, c0.cout, &c1);
    if (rc) return rc;
    LIP_REQUIRE(c1.stride == 1, "lip_model_create: second conv of the block at %d must have stride 1", i);
    int j = i + 6;
    bool proj = false;
    ConvBN pj;
    if (is(j, LIP_OP_RES_CONV2D)) {
      LIP_REQUIRE(is(j + 1, LIP_OP_RES_BATCHNORM), "lip_model_create: RES_CONV2D at %d must be followed by RES_BATCHNORM", j);
      rc = conv_bn(j, j + 1, H, W, C, &pj);
      if (rc) return rc;
      LIP_REQUIRE(pj.Ho == c1.Ho && pj.Wo == c1.Wo && pj.cout == c1.cout, "lip_model_create: shortcut at %d does not match the branch shape", j);
      pj.relu = 0; pj.src = x; pj.dst = r; pj.skip = -1; pj.accumulate = 0;
      proj = true;
      j += 2;
    } else {
      LIP_REQUIRE(c1.Ho == H && c1.Wo == W && c1.cout == C, "lip_model_create: block at %d changes shape but has no shortcut conv", i);
    }
    LIP_REQUIRE(is(j, LIP_OP_RES_ADD) && is(j + 1, LIP_OP_RELU), "lip_model_create: block at %d must end with RES_ADD, RELU", i);
    c1.relu = 1; c1.src = y; c1.dst = o; c1.skip = proj ? r : x; c1.accumulate = 0;
    m->RB.push_back(c0);
    if (proj) m->RB.push_back(pj);
    m->RB.push_back(c1);
    H = c1.Ho; W = c1.Wo; C = c1.cout;
    x = o;
    i = j + 2;
  }
  LIP_REQUIRE(is(i, LIP_OP_GLOBAL_MEAN) && is(i + 1, LIP_OP_DENSE) && i + 2 == n,
              "lip_model_create: a residual program ends with GLOBAL_MEAN, DENSE (op %d of %d)", i, n);
  const lip_layer_desc& d = L[i + 1];
  LIP_REQUIRE(d.in_features == C && d.out_features > 0 && d.bias_offset >= 0 && d.kernel_offset >= 0 &&
                  d.bias_offset + d.out_features <= num_params && d.kernel_offset + (int64_t)C * d.out_features <= num_params,
              "lip_model_create: DENSE head: in_features %d != %d channels (or bad offset)", d.in_features, C);
  counted += (int64_t)C * d.out_features + d.out_features;
  LIP_REQUIRE(counted == num_params, "lip_model_create: layers hold %lld parameters but num_params = %lld", (long long)counted,
              (long long)num_params);
  m->rn_last_slot = x; m->rn_H = H; m->rn_W = W; m->rn_C = C;
  m->rn_dense_boff = d.bias_offset; m->rn_dense_woff = d.kernel_offset;
  m->rn_nstats = stats;
  m->rn_slot_elems = max_elems;
  m->K = d.out_features;
  return LIP_OK;
}

int resnet_bind(lip_model* m, const float* theta, const float* Z, int64_t M, cudaStream_t st) {
  LIP_REQUIRE(m->rn_stats != nullptr, "lip_model_bind: BatchNorm statistics missing (call lip_model_set_bn_stats first)");
  float* act[4] = {nullptr, nullptr, nullptr, nullptr};
  float* tmp = nullptr;
  size_t max_raw = 0;
  for (const ConvBN& u : m->RB) {
    const size_t r = (size_t)M * u.P() * u.cout;
    max_raw = r > max_raw ? r : max_raw;
  }
  auto cleanup = [&]() { for (auto& a : act) if (a) cudaFree(a); if (tmp) cudaFree(tmp); };
  for (auto& a : act) {
    if (cudaMalloc(&a, sizeof(float) * (size_t)M * m->rn_slot_elems + 256) != cudaSuccess) {
      cleanup(); set_error("lip_model_bind: out of device memory (activations)"); return LIP_ERR_CUDA;
    }
  }
  if (cudaMalloc(&tmp, sizeof(float) * max_raw + 256) != cudaSuccess) {
    cleanup(); set_error("lip_model_bind: out of device memory"); return LIP_ERR_CUDA;
  }
  int rc = LIP_OK;
  bool any_tc = false;
  if (m->use_tc == 1 && !tc_available()) {
    cleanup(); set_error("lip_model_bind: tensor path requested but the device is not sm_100 / TMA encode unavailable");
    return LIP_ERR_UNSUPPORTED;
  }
  for (ConvBN& u : m->RB) {
    const int64_t R = M * u.P(), Kc = u.Kc();
    const size_t xin_elems = (size_t)M * u.Hi * u.Wi * u.cin;
    if (cudaMalloc(&u.Xin, sizeof(float) * xin_elems + 256) != cudaSuccess ||
        cudaMalloc(&u.Wt, sizeof(float) * (size_t)Kc * u.cout + 256) != cudaSuccess ||
        cudaMalloc(&u.xhat, sizeof(float) * (size_t)R * u.cout + 256) != cudaSuccess ||
        cudaMalloc(&u.g, sizeof(float) * u.cout) != cudaSuccess ||
        (u.relu && cudaMalloc(&u.mask, sizeof(float) * (size_t)R * u.cout + 256) != cudaSuccess)) {
      cleanup(); set_error("lip_model_bind: out of device memory (conv cache)"); return LIP_ERR_CUDA;
    }
    const float* x = u.src == -2 ? Z : act[u.src];
    if (cudaMemcpyAsync(u.Xin, x, sizeof(float) * xin_elems, cudaMemcpyDeviceToDevice, st) != cudaSuccess) {
      cleanup(); set_error("lip_model_bind: activation copy failed"); return LIP_ERR_CUDA;
    }
    conv_wt_kernel<<<(unsigned)ceil_div(Kc * u.cout, 256), 256, 0, st>>>(theta + u.woff, u.Wt, u.kh * u.kw, u.cin, u.cout);
    count_launch();
    u.tc = m->use_tc != 0 && u.src != -2 && u.pad_h == u.pad_w &&
           conv_tc_supported(u.Hi, u.Wi, u.cin, u.cout, u.kh, u.kw, u.stride, u.pad_h);
    if (u.tc) {   // TF32 (hi, lo) splits of the cached image and of both kernel layouts
      const size_t wk = (size_t)Kc * u.cout;
      if (cudaMalloc(&u.Xh, sizeof(float) * xin_elems + 256) != cudaSuccess ||
          cudaMalloc(&u.Xl, sizeof(float) * xin_elems + 256) != cudaSuccess ||
          cudaMalloc(&u.Wh, sizeof(float) * wk + 256) != cudaSuccess || cudaMalloc(&u.Wl, sizeof(float) * wk + 256) != cudaSuccess ||
          cudaMalloc(&u.Wth, sizeof(float) * wk + 256) != cudaSuccess || cudaMalloc(&u.Wtl, sizeof(float) * wk + 256) != cudaSuccess) {
        cleanup(); set_error("lip_model_bind: out of device memory (tensor-path conv cache)"); return LIP_ERR_CUDA;
      }
      rc = tf32_split(u.Xin, (int64_t)xin_elems, u.Xh, u.Xl, (int64_t)xin_elems, 1, (int64_t)xin_elems, st);
      if (!rc) rc = tf32_split(theta + u.woff, (int64_t)wk, u.Wh, u.Wl, (int64_t)wk, 1, (int64_t)wk, st);
      if (!rc) rc = tf32_split(u.Wt, (int64_t)wk, u.Wth, u.Wtl, (int64_t)wk, 1, (int64_t)wk, st);
      if (rc) break;
      any_tc = true;
    }
    GemmProblem p;
    p.M = R; p.N = u.cout; p.K = Kc; p.batch = 1;
    p.A1.ptr = u.Xin; p.A1.sz = 0; p.A1.conv = u.gather(1);
    p.B1 = {theta + u.woff, 0, u.cout, 1};
    p.C = tmp; p.c_sz = 0; p.c_sm = u.cout;
    rc = gemm_simt(p, st);
    if (rc) break;
    const long long total = R * (long long)u.cout;
    bn_fwd_kernel<<<ew_grid(total), 256, 0, st>>>(tmp, m->rn_stats + u.stats_off, m->rn_stats + u.stats_off + u.cout,
                                                  theta + u.scale_off, theta + u.beta_off, u.skip >= 0 ? act[u.skip] : nullptr,
                                                  u.xhat, u.mask, u.g, act[u.dst], total, u.cout, u.relu);
    count_launch();
    if (cudaGetLastError() != cudaSuccess) { rc = LIP_ERR_CUDA; set_error("lip_model_bind: bn_fwd launch failed"); break; }
  }
  if (!rc) {
    const int HW = m->rn_H * m->rn_W, C = m->rn_C;
    if (cudaMalloc(&m->rn_mean_act, sizeof(float) * (size_t)M * C + 256) != cudaSuccess ||
        cudaMalloc(&m->logits, sizeof(float) * (size_t)M * m->K) != cudaSuccess ||
        cudaMalloc(&m->P, sizeof(float) * (size_t)M * m->K) != cudaSuccess ||
        cudaMalloc(&m->S, sizeof(float) * (size_t)M * m->K) != cudaSuccess) {
      cleanup(); set_error("lip_model_bind: out of device memory (head)"); return LIP_ERR_CUDA;
    }
    const long long total = (long long)M * C;
    global_mean_kernel<<<ew_grid(total), 256, 0, st>>>(act[m->rn_last_slot], nullptr, m->rn_mean_act, total, HW, C);
    count_launch();
    GemmProblem p;
    p.M = M; p.N = m->K; p.K = C; p.batch = 1;
    p.A1 = {m->rn_mean_act, 0, C, 1};
    p.B1 = {theta + m->rn_dense_woff, 0, m->K, 1};
    p.C = m->logits; p.c_sz = 0; p.c_sm = m->K;
    p.epi.bias = theta + m->rn_dense_boff; p.epi.bias_sz = 0;
    rc = gemm_simt(p, st);
    if (!rc && m->model_type == LIP_CLASSIFIER) rc = launch_softmax(m->logits, m->P, m->S, M, m->K, st);
  }
  // the temporaries are read by work queued on `st`: free them only after it has drained (bind is not a hot call)
  cudaStreamSynchronize(st);
  cleanup();
  if (rc) return rc;
  if (any_tc && !m->lo_nz && cudaMalloc(&m->lo_nz, sizeof(int) * m->RB.size()) != cudaSuccess) {
    set_error("lip_model_bind: out of device memory (flags)"); return LIP_ERR_CUDA;
  }
  m->tc_on = any_tc;
  static const int no_fuse = getenv("LIP_CONV_TC_FUSE") ? (atoi(getenv("LIP_CONV_TC_FUSE")) == 0) : 0;
  m->rn_pairs = any_tc && !no_fuse;
  for (const ConvBN& u : m->RB) if (u.src != -2 && !u.tc) m->rn_pairs = false;
  m->tc_layer.assign(m->RB.size(), 0);
  for (size_t i = 0; i < m->RB.size(); ++i) m->tc_layer[i] = m->RB[i].tc ? 1 : 0;
  m->bound = true;
  return LIP_OK;
}

size_t resnet_ws_bytes(const lip_model* m, int64_t B) {
  const RnSizes z = rn_sizes(m, B);
  return (4 * z.slot + z.raw + z.col + z.head + z.dl + z.raw_lo + 2 * z.img + 2 * z.vsplit + 4 * z.slot_lo) * sizeof(float) + 512;
}

int resnet_ggn_vp(lip_model* m, const float* V, float* out, int64_t B, float recal, float alpha, void* ws, size_t bytes,
                  cudaStream_t st) {
  RnWs w;
  int rc = rn_carve(m, B, ws, bytes, &w);
  if (rc) return rc;
  rc = rn_jvp_sweep(m, V, B, w, w.dl, st);
  if (rc) return rc;
  if (m->model_type == LIP_CLASSIFIER) {
    rc = launch_factor(w.dl, w.dl, m, B, 0, 1.f, st);
    if (rc) return rc;
  }
  return rn_vjp_sweep(m, w.dl, B, w, out, recal, alpha != 0.f ? V : nullptr, alpha, st);
}

int resnet_wt_apply(lip_model* m, const float* V, float* out, int64_t B, float scale, int32_t factor, void* ws, size_t bytes,
                    cudaStream_t st) {
  RnWs w;
  int rc = rn_carve(m, B, ws, bytes, &w);
  if (rc) return rc;
  rc = rn_jvp_sweep(m, V, B, w, out, st);
  if (rc) return rc;
  float s = scale;
  if (factor == LIP_FACTOR_SQRT && m->model_type == LIP_REGRESSOR) s *= expf(-0.5f * m->logvar);
  if (factor == LIP_FACTOR_SQRT && m->model_type == LIP_CLASSIFIER) return launch_factor(out, out, m, B, 1, s, st);
  if (s != 1.f) return launch_scale_copy(out, out, B * m->M * m->K, s, st);
  return LIP_OK;
}

// ---- lip_zgrad for residual programs ---------------------------------------------------------------------------------------------
// The network is piecewise linear in its activations (relu, BatchNorm in eval mode, convs, adds, mean), so the recurrences of
// lip_zgrad.cu need no second derivative and no stored tangent; what is new against the plain conv stage programs (lip_cnn.cu) is
//   * the BatchNorm SCALE tangent: dY contains xhat(Z) * dscale[b], and xhat = rstd * (conv(x, W) - mean) depends on Z, which adds
//     rstd * dscale[b] * ebar to the cotangent that travels back through the shared kernel;
//   * skip connections: both cotangent families (e = d s / d tangent-activation, q = d s / d activation) follow the slot rotation
//     of rn_vjp_sweep (block input, branch, shortcut, output).
// Per unit, in reverse order:   ebar = mask * e_out,  qbar = mask * q_out,  (e, q)[skip] <- (ebar, qbar)
//     e[src] (+)= convT(g * ebar, W)
//     q[src] (+)= convT(g * qbar + rstd * dscale[b] * ebar, W) + convT(g * ebar, dW[b])          (one dual-K implicit GEMM)
// and at the stem dZ[b] = q[input].  All GEMMs are the fp32 SIMT implicit GEMMs (transposed-conv gather, ConvGather mode 3).
struct RnZSizes { size_t inner, slot, raw, wt, head, small, fin, sum, total; };

static RnZSizes rn_zsizes(const lip_model* m, int nseg, int64_t B) {
  RnZSizes z{};
  const RnSizes r = rn_sizes(m, B);
  z.inner = align_up(resnet_ws_bytes(m, B), 256);
  z.slot = r.slot;
  z.raw = r.raw;
  size_t wt = 0;
  for (const ConvBN& u : m->RB) { const size_t v = (size_t)B * u.Kc() * u.cout; wt = v > wt ? v : wt; }
  z.wt = align_up(wt, 64);
  z.head = align_up((size_t)B * m->M * m->rn_C, 64);
  z.small = align_up((size_t)nseg * B * m->M * m->K, 64);
  z.fin = align_up((size_t)nseg * B * m->M * (size_t)m->in_h * m->in_w * m->in_c, 64);
  z.sum = align_up((size_t)B * m->M * (size_t)m->in_h * m->in_w * m->in_c, 64);
  z.total = z.inner + (4 * z.slot + z.raw + z.wt + z.head + 3 * z.small + z.fin + z.sum) * sizeof(float) + 1024;
  return z;
}

size_t resnet_zgrad_ws_bytes(const lip_model* m, int32_t mode, int64_t B) {
  return rn_zsizes(m, mode == LIP_ZGRAD_GGN ? 2 : 1, B).total;
}

int resnet_zgrad(lip_model* m, int32_t mode, const float* X1, const float* X2, float* out, int64_t B, float scale, int32_t per_probe,
                 void* ws, size_t bytes, cudaStream_t st) {
  const int nseg = mode == LIP_ZGRAD_GGN ? 2 : 1;
  const RnZSizes z = rn_zsizes(m, nseg, B);
  if (!ws || bytes < z.total) {
    set_error("lip_zgrad: workspace too small: need %zu bytes, got %zu", z.total, bytes);
    return LIP_ERR_WORKSPACE;
  }
  RnWs w;
  int rc = rn_carve(m, B, ws, z.inner, &w);
  if (rc) return rc;
  float* p = (float*)align_up((uintptr_t)ws + z.inner, 256);
  float* Qs[4];
  for (int i = 0; i < 4; ++i) { Qs[i] = p; p += z.slot; }
  float* rawQ = p; p += z.raw;
  float* wtb = p; p += z.wt;
  float* qhead = p; p += z.head;
  float* dl = p; p += z.small;
  float* Cc = p; p += z.small;
  float* Gf = p; p += z.small;
  float* fin = p; p += z.fin;
  float* sum = p;
  float** Es = w.slot;
  float* rawE = w.raw;
  const float* Vseg[2] = {X1, mode == LIP_ZGRAD_GGN ? X2 : nullptr};
  const int64_t M = m->M;
  const int64_t small_seg = B * M * m->K;
  const int HW = m->rn_H * m->rn_W, C = m->rn_C;
  const int64_t img = (int64_t)m->in_h * m->in_w * m->in_c;

  for (int sg = 0; sg < nseg; ++sg) {   // forward tangent pass: only the tangent logits are needed
    rc = rn_jvp_sweep(m, Vseg[sg], B, w, dl + sg * small_seg, st);
    if (rc) return rc;
  }
  rc = launch_zgrad_rows(mode, m, dl, X2, Cc, Gf, B, scale, st);
  if (rc) return rc;

  for (int sg = 0; sg < nseg; ++sg) {
    const float* V = Vseg[sg];
    const float* c_rows = Cc + sg * small_seg;
    const float* g_rows = Gf + sg * small_seg;
    {  // head: e_mean = c Wd^T,  q_mean = c dWd[b]^T + g Wd^T, then back through the global mean
      GemmProblem e;
      e.M = M; e.N = C; e.K = m->K; e.batch = B;
      e.A1 = {c_rows, M * (int64_t)m->K, m->K, 1};
      e.B1 = {m->theta + m->rn_dense_woff, 0, 1, m->K};
      e.C = w.head; e.c_sz = M * (int64_t)C; e.c_sm = C;
      rc = gemm_simt(e, st);
      if (rc) return rc;
      GemmProblem q;
      q.M = M; q.N = C; q.K = m->K; q.batch = B;
      q.A1 = {c_rows, M * (int64_t)m->K, m->K, 1};
      q.B1 = {V + m->rn_dense_woff, m->D, 1, m->K};
      q.A2 = {g_rows, M * (int64_t)m->K, m->K, 1};
      q.B2 = {m->theta + m->rn_dense_woff, 0, 1, m->K};
      q.K2 = m->K;
      q.C = qhead; q.c_sz = M * (int64_t)C; q.c_sm = C;
      rc = gemm_simt(q, st);
      if (rc) return rc;
      const long long total = (long long)B * M * HW * C;
      global_mean_bwd_kernel<<<ew_grid(total), 256, 0, st>>>(w.head, Es[m->rn_last_slot], total, HW, C);
      LIP_LAUNCH_CHECK();
      global_mean_bwd_kernel<<<ew_grid(total), 256, 0, st>>>(qhead, Qs[m->rn_last_slot], total, HW, C);
      LIP_LAUNCH_CHECK();
    }
    for (int i = (int)m->RB.size() - 1; i >= 0; --i) {
      const ConvBN& u = m->RB[i];
      const int64_t R = M * u.P(), Kc = u.Kc();
      const long long per_z = R * (long long)u.cout, total = per_z * B;
      rn_zgrad_bn_kernel<<<ew_grid(total), 256, 0, st>>>(Es[u.dst], Qs[u.dst], u.mask, u.g, m->rn_stats + u.stats_off + u.cout,
                                                         V + u.scale_off, m->D, u.skip >= 0 ? Es[u.skip] : nullptr,
                                                         u.skip >= 0 ? Qs[u.skip] : nullptr, rawE, rawQ, total, per_z, u.cout);
      LIP_LAUNCH_CHECK();
      {
        const int tot = (int)(Kc * u.cout);
        dim3 grid((unsigned)std::min<int64_t>(ceil_div(tot, 256), 256), (unsigned)B);
        conv_wt_batched_kernel<<<grid, 256, 0, st>>>(V + u.woff, m->D, wtb, u.kh * u.kw, u.cin, u.cout);
        LIP_LAUNCH_CHECK();
      }
      const int64_t in_elems = M * (int64_t)u.Hi * u.Wi * u.cin;
      if (u.src != -2) {   // e[src] (+)= convT(rawE, W)
        GemmProblem e;
        e.M = M * (int64_t)u.Hi * u.Wi; e.N = u.cin; e.K = (int64_t)u.kh * u.kw * u.cout; e.batch = B;
        e.A1.ptr = rawE; e.A1.sz = per_z; e.A1.conv = u.gather(3);
        e.B1 = {u.Wt, 0, u.cin, 1};
        e.C = Es[u.src]; e.c_sz = in_elems; e.c_sm = u.cin;
        if (u.accumulate) { e.epi.add = e.C; e.epi.add_sz = e.c_sz; e.epi.add_scale = 1.f; }
        rc = gemm_simt(e, st);
        if (rc) return rc;
      }
      GemmProblem q;       // q[src] (+)= convT(rawQ, W) + convT(rawE, dW[b])
      q.M = M * (int64_t)u.Hi * u.Wi; q.N = u.cin; q.K = (int64_t)u.kh * u.kw * u.cout; q.batch = B;
      q.A1.ptr = rawQ; q.A1.sz = per_z; q.A1.conv = u.gather(3);
      q.B1 = {u.Wt, 0, u.cin, 1};
      q.A2.ptr = rawE; q.A2.sz = per_z; q.A2.conv = u.gather(3);
      q.B2 = {wtb, Kc * (int64_t)u.cout, u.cin, 1};
      q.K2 = q.K;
      q.C = (u.src == -2) ? fin + (int64_t)sg * B * M * img : Qs[u.src]; q.c_sz = in_elems; q.c_sm = u.cin;
      if (u.src != -2 && u.accumulate) { q.epi.add = q.C; q.epi.add_sz = q.c_sz; q.epi.add_scale = 1.f; }
      rc = gemm_simt(q, st);
      if (rc) return rc;
    }
  }
  // sum over probes (and, in GGN mode, over the two tangent families)
  const int64_t per0 = M * img;
  if (!per_probe) return launch_batch_sum(fin, out, per0, (int64_t)nseg * B, st);
  if (nseg == 2) return launch_batch_sum(fin, out, B * per0, 2, st);
  (void)sum;
  return launch_scale_copy(fin, out, B * per0, 1.f, st);
}

int resnet_w_apply(lip_model* m, const float* U, float* out, int64_t B, float scale, int32_t factor, const float* add,
                   float add_scale, void* ws, size_t bytes, cudaStream_t st) {
  RnWs w;
  int rc = rn_carve(m, B, ws, bytes, &w);
  if (rc) return rc;
  float s = scale;
  if (factor == LIP_FACTOR_SQRT && m->model_type == LIP_CLASSIFIER) {
    rc = launch_factor(U, w.dl, m, B, 2, 1.f, st);
  } else {
    if (factor == LIP_FACTOR_SQRT) s *= expf(-0.5f * m->logvar);
    rc = launch_scale_copy(U, w.dl, B * m->M * m->K, 1.f, st);
  }
  if (rc) return rc;
  return rn_vjp_sweep(m, w.dl, B, w, out, s, add, add_scale, st);
}

}  // namespace lip
