// lip_cnn.cu — probe-batched GGN / W / W^T operators for conv stage programs (LeNet5, src/scalemodels.py:11-49).
//
// Reference semantics are the same as lip_model.cu (src/ggn.py:9-146): only the model differs.  A CONV2D stage is run as
// a GEMM over im2col patches, so every stage (conv or dense) has the MLP form
//   JVP  stage s:  dH_s[b] = Aop_s . dW_s[b] + col(T_s[b]) . W_s + db_s[b];   dA_s = phi'_s * dH_s;   T_{s+1} = avgpool(dA_s)
//   VJP  stage s:  gW_s[b] = Aop_s^T . D_s[b];  gb_s[b] = colsum D_s[b];  G = D_s[b] . W_s^T;
//                  T = col2im(G);  D_{s-1}[b] = phi'_{s-1} * unpool(T) / 4
// with rows = (point, output pixel), Aop_s = the im2col patches of the cached forward activations (bound once) and
// col(.) = im2col of the per-probe tangent image.  Patches / tangent images are NHWC, kernels HWIO, so the patch column
// order (dy, dx, ci) is exactly the row order of the flat flax kernel [kh, kw, cin, cout] and dW_s[b] / gW_s[b] are read
// and written IN PLACE in the caller's [B, D] blocks, as in the MLP path.
// LeNet's channel counts (1, 6, 16) are hostile to 128-wide MMA tiles: all GEMMs here run on the exact-fp32 SIMT kernels
// (gemm_simt: 128x128 tiles, or the skinny N <= 16 variant).
#include <new>

#include "lip_model.cuh"

using namespace lip;

void lip_model::free_cnn_cache() {
  for (auto& s : CS) {
    if (s.Aop) cudaFree(s.Aop);
    if (s.dphi) cudaFree(s.dphi);
    if (s.Xin) cudaFree(s.Xin);
    s.Aop = s.dphi = s.Xin = nullptr;
  }
  if (cnn_tmp_out) cudaFree(cnn_tmp_out);
  if (cnn_tmp_x) cudaFree(cnn_tmp_x);
  cnn_tmp_out = cnn_tmp_x = nullptr;
}

namespace {

// out[r][(dy*kw + dx)*C + c] = in[mz][y*stride + dy - pad_h][x*stride + dx - pad_w][c]   (0 outside), r = (mz*Ho + y)*Wo + x
__global__ void im2col_kernel(const float* __restrict__ in, float* __restrict__ out, long long total, int Hi, int Wi, int C,
                              int pad_h, int pad_w, int stride, int kh, int kw, int Ho, int Wo) {
  for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total; idx += (long long)gridDim.x * blockDim.x) {
    long long t = idx;
    const int c = (int)(t % C); t /= C;
    const int dx = (int)(t % kw); t /= kw;
    const int dy = (int)(t % kh); t /= kh;
    const int x = (int)(t % Wo); t /= Wo;
    const int y = (int)(t % Ho); t /= Ho;
    const int yi = y * stride + dy - pad_h, xi = x * stride + dx - pad_w;
    float v = 0.f;
    if (yi >= 0 && yi < Hi && xi >= 0 && xi < Wi) v = __ldg(in + ((t * Hi + yi) * Wi + xi) * C + c);
    out[idx] = v;
  }
}

// Same result, one image per blockIdx.y step: the element index inside an image fits 32 bits, so the (pixel, tap, channel)
// decomposition is four multiply-shift divisions (FastDiv) instead of five 64-bit divisions per element - the 64-bit form was
// integer-bound at 0.9 TB/s of patch writes (profiles/r01_launches_lenet5_summary.txt).
__global__ void __launch_bounds__(256) im2col_img_kernel(const float* __restrict__ in, float* __restrict__ out, long long MZ, int per_img,
                                                         int Hi, int Wi, int C, int pad_h, int pad_w, int stride, int kw, int Ho, int Wo,
                                                         lip::FastDiv dKc, lip::FastDiv dRow, lip::FastDiv dC, lip::FastDiv dWo) {
  (void)Ho;
  for (long long mz = blockIdx.y; mz < MZ; mz += gridDim.y) {
    const float* img = in + mz * (long long)Hi * Wi * C;
    float* o = out + mz * (long long)per_img;
    for (uint32_t e = blockIdx.x * blockDim.x + threadIdx.x; e < (uint32_t)per_img; e += gridDim.x * blockDim.x) {
      uint32_t r, kc, dy, j, dx, c, y, x;
      dKc.divmod(e, r, kc);        // output pixel, patch column
      dRow.divmod(kc, dy, j);      // tap row, (dx, c)
      dC.divmod(j, dx, c);
      dWo.divmod(r, y, x);
      const int yi = (int)y * stride + (int)dy - pad_h, xi = (int)x * stride + (int)dx - pad_w;
      float v = 0.f;
      if (yi >= 0 && yi < Hi && xi >= 0 && xi < Wi) v = __ldg(img + ((long long)yi * Wi + xi) * C + c);
      o[e] = v;
    }
  }
}

// tin[mz][yi][xi][c] (+)= sum over (dy, dx) with y*stride = yi + pad_h - dy, x*stride = xi + pad_w - dx, 0 <= y < Ho, 0 <= x < Wo
//                         of col[(mz*Ho + y)*Wo + x][(dy*kw + dx)*C + c]          (gather form: deterministic, no atomics)
__global__ void col2im_kernel(const float* __restrict__ col, float* __restrict__ tin, long long total, int Hi, int Wi, int C,
                              int pad_h, int pad_w, int stride, int kh, int kw, int Ho, int Wo, int accumulate) {
  const int Kc = kh * kw * C;
  for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total; idx += (long long)gridDim.x * blockDim.x) {
    long long t = idx;
    const int c = (int)(t % C); t /= C;
    const int xi = (int)(t % Wi); t /= Wi;
    const int yi = (int)(t % Hi); t /= Hi;
    float acc = 0.f;
    for (int dy = 0; dy < kh; ++dy) {
      const int ys = yi + pad_h - dy;
      if (ys < 0 || ys % stride != 0) continue;
      const int y = ys / stride;
      if (y >= Ho) continue;
      for (int dx = 0; dx < kw; ++dx) {
        const int xs = xi + pad_w - dx;
        if (xs < 0 || xs % stride != 0) continue;
        const int x = xs / stride;
        if (x >= Wo) continue;
        acc += __ldg(col + ((t * Ho + y) * Wo + x) * Kc + (dy * kw + dx) * C + c);
      }
    }
    tin[idx] = accumulate ? tin[idx] + acc : acc;
  }
}

// stride-1 form with one image per blockIdx.y step and multiply-shift index decomposition (no division in the tap loop)
__global__ void __launch_bounds__(256) col2im_s1_kernel(const float* __restrict__ col, float* __restrict__ tin, long long MZ, int per_img,
                                                        int Hi, int Wi, int C, int pad_h, int pad_w, int kh, int kw, int Ho, int Wo,
                                                        int accumulate, lip::FastDiv dC, lip::FastDiv dWi) {
  (void)Hi;
  const int Kc = kh * kw * C;
  for (long long mz = blockIdx.y; mz < MZ; mz += gridDim.y) {
    const float* cimg = col + mz * (long long)Ho * Wo * Kc;
    float* o = tin + mz * (long long)per_img;
    for (uint32_t e = blockIdx.x * blockDim.x + threadIdx.x; e < (uint32_t)per_img; e += gridDim.x * blockDim.x) {
      uint32_t t, c, yi, xi;
      dC.divmod(e, t, c);
      dWi.divmod(t, yi, xi);
      float acc = 0.f;
      const int dy0 = (int)yi + pad_h - (Ho - 1) > 0 ? (int)yi + pad_h - (Ho - 1) : 0;
      const int dy1 = (int)yi + pad_h < kh - 1 ? (int)yi + pad_h : kh - 1;
      const int dx0 = (int)xi + pad_w - (Wo - 1) > 0 ? (int)xi + pad_w - (Wo - 1) : 0;
      const int dx1 = (int)xi + pad_w < kw - 1 ? (int)xi + pad_w : kw - 1;
      for (int dy = dy0; dy <= dy1; ++dy) {
        const int y = (int)yi + pad_h - dy;
        for (int dx = dx0; dx <= dx1; ++dx) {
          const int x = (int)xi + pad_w - dx;
          acc += __ldg(cimg + ((long long)y * Wo + x) * Kc + (dy * kw + dx) * C + c);
        }
      }
      o[e] = accumulate ? o[e] + acc : acc;
    }
  }
}

// out[mz][yp][xp][c] = mean of the 2x2 window of in[mz][.][.][c]   (in: Ho x Wo, out: Ho/2 x Wo/2)
__global__ void avgpool2_kernel(const float* __restrict__ in, float* __restrict__ out, long long total, int Ho, int Wo, int C) {
  const int Hp = Ho / 2, Wp = Wo / 2;
  for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total; idx += (long long)gridDim.x * blockDim.x) {
    long long t = idx;
    const int c = (int)(t % C); t /= C;
    const int xp = (int)(t % Wp); t /= Wp;
    const int yp = (int)(t % Hp); t /= Hp;
    const float* b = in + ((t * Ho + 2 * yp) * Wo + 2 * xp) * C + c;
    out[idx] = 0.25f * ((b[0] + b[C]) + (b[(long long)Wo * C] + b[(long long)Wo * C + C]));
  }
}

// d[z][r][c] = dphi[r][c] * (pool ? 0.25 * tin[z][m][y/2][x/2][c] : tin[z][r][c]),  r = (m*Ho + y)*Wo + x, per_z = M*Ho*Wo*C
__global__ void unpool_mask_kernel(const float* __restrict__ tin, const float* __restrict__ dphi, float* __restrict__ d,
                                   long long total, long long per_z, int Ho, int Wo, int C, int pool) {
  const int Hp = Ho / 2, Wp = Wo / 2;
  for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total; idx += (long long)gridDim.x * blockDim.x) {
    const long long z = idx / per_z, rem = idx % per_z;
    float v;
    if (pool) {
      long long t = rem;
      const int c = (int)(t % C); t /= C;
      const int x = (int)(t % Wo); t /= Wo;
      const int y = (int)(t % Ho); t /= Ho;     // t = m
      v = 0.25f * __ldg(tin + (((z * (per_z / ((long long)Ho * Wo * C)) + t) * Hp + (y >> 1)) * Wp + (x >> 1)) * C + c);
    } else {
      v = __ldg(tin + idx);
    }
    d[idx] = v * __ldg(dphi + rem);
  }
}

// 32-bit form: blockIdx.y strides over the batch entries z, the element index inside one entry (per_z < 2^31) is decomposed with
// multiply-shift divisions
__global__ void __launch_bounds__(256) unpool_mask_z_kernel(const float* __restrict__ tin, const float* __restrict__ dphi,
                                                            float* __restrict__ d, long long B, int per_z, long long tin_per_z,
                                                            int Ho, int Wo, int C, lip::FastDiv dC, lip::FastDiv dWo, lip::FastDiv dHo) {
  const int Hp = Ho / 2, Wp = Wo / 2;
  for (long long z = blockIdx.y; z < B; z += gridDim.y) {
    const float* ti = tin + z * tin_per_z;
    float* dz = d + z * (long long)per_z;
    for (uint32_t e = blockIdx.x * blockDim.x + threadIdx.x; e < (uint32_t)per_z; e += gridDim.x * blockDim.x) {
      uint32_t t, c, x, y, m;
      dC.divmod(e, t, c);
      dWo.divmod(t, t, x);
      dHo.divmod(t, m, y);
      const float v = 0.25f * __ldg(ti + (((long long)m * Hp + (y >> 1)) * Wp + (x >> 1)) * C + c);
      dz[e] = v * __ldg(dphi + e);
    }
  }
}

inline unsigned ew_grid(long long total) {
  long long g = (total + 255) / 256;
  if (g > 148ll * 32) g = 148ll * 32;
  if (g < 1) g = 1;
  return (unsigned)g;
}
}  // namespace

namespace lip {
int im2col(const float* in, float* out, int64_t MZ, int Hi, int Wi, int C, int pad_h, int pad_w, int stride, int kh, int kw,
           int Ho, int Wo, cudaStream_t st) {
  const long long total = (long long)MZ * Ho * Wo * kh * kw * C;
  const long long per_img = (long long)Ho * Wo * kh * kw * C;
  if (per_img < (1LL << 31) && MZ > 0) {
    const int Kc = kh * kw * C;
    long long gx = (per_img + 255) / 256;
    if (gx > 64) gx = 64;
    long long gy = MZ < 148 * 32 / gx + 1 ? MZ : 148 * 32 / gx + 1;
    if (gy > 65535) gy = 65535;
    dim3 grid((unsigned)gx, (unsigned)gy);
    im2col_img_kernel<<<grid, 256, 0, st>>>(in, out, MZ, (int)per_img, Hi, Wi, C, pad_h, pad_w, stride, kw, Ho, Wo, FastDiv((uint32_t)Kc),
                                            FastDiv((uint32_t)(kw * C)), FastDiv((uint32_t)C), FastDiv((uint32_t)Wo));
    LIP_LAUNCH_CHECK();
    return LIP_OK;
  }
  im2col_kernel<<<ew_grid(total), 256, 0, st>>>(in, out, total, Hi, Wi, C, pad_h, pad_w, stride, kh, kw, Ho, Wo);
  LIP_LAUNCH_CHECK();
  return LIP_OK;
}
int col2im(const float* col, float* tin, int64_t MZ, int Hi, int Wi, int C, int pad_h, int pad_w, int stride, int kh, int kw,
           int Ho, int Wo, int accumulate, cudaStream_t st) {
  const long long total = (long long)MZ * Hi * Wi * C;
  const long long per_img = (long long)Hi * Wi * C;
  if (stride == 1 && per_img < (1LL << 31) && MZ > 0) {
    long long gx = (per_img + 255) / 256;
    if (gx > 64) gx = 64;
    long long gy = MZ < 148 * 32 / gx + 1 ? MZ : 148 * 32 / gx + 1;
    if (gy > 65535) gy = 65535;
    dim3 grid((unsigned)gx, (unsigned)gy);
    col2im_s1_kernel<<<grid, 256, 0, st>>>(col, tin, MZ, (int)per_img, Hi, Wi, C, pad_h, pad_w, kh, kw, Ho, Wo, accumulate,
                                           FastDiv((uint32_t)C), FastDiv((uint32_t)Wi));
    LIP_LAUNCH_CHECK();
    return LIP_OK;
  }
  col2im_kernel<<<ew_grid(total), 256, 0, st>>>(col, tin, total, Hi, Wi, C, pad_h, pad_w, stride, kh, kw, Ho, Wo, accumulate);
  LIP_LAUNCH_CHECK();
  return LIP_OK;
}
}  // namespace lip

namespace {

int launch_im2col(const float* in, float* out, int64_t MZ, const ConvStage& s, cudaStream_t st) {
  return im2col(in, out, MZ, s.Hi, s.Wi, s.cin, s.pad, s.pad, 1, s.kh, s.kw, s.Ho, s.Wo, st);
}
int launch_col2im(const float* col, float* tin, int64_t MZ, const ConvStage& s, cudaStream_t st) {
  return col2im(col, tin, MZ, s.Hi, s.Wi, s.cin, s.pad, s.pad, 1, s.kh, s.kw, s.Ho, s.Wo, 0, st);
}
int launch_avgpool(const float* in, float* out, int64_t MZ, const ConvStage& s, cudaStream_t st) {
  const long long total = (long long)MZ * s.Hp * s.Wp * s.cout;
  avgpool2_kernel<<<ew_grid(total), 256, 0, st>>>(in, out, total, s.Ho, s.Wo, s.cout);
  LIP_LAUNCH_CHECK();
  return LIP_OK;
}
// d (gradient w.r.t. the pre-activation of stage s, [B, M*P, cout]) from tin (gradient w.r.t. the stage's output)
int launch_unpool_mask(const float* tin, float* d, int64_t B, int64_t M, const ConvStage& s, cudaStream_t st) {
  const long long per_z = (long long)M * s.P * s.cout, total = per_z * B;
  if (s.pool && per_z < (1LL << 31)) {
    long long gx = (per_z + 255) / 256;
    if (gx > 296) gx = 296;
    long long gy = B < 148 * 32 / gx + 1 ? B : 148 * 32 / gx + 1;
    if (gy > 65535) gy = 65535;
    dim3 grid((unsigned)gx, (unsigned)gy);
    unpool_mask_z_kernel<<<grid, 256, 0, st>>>(tin, s.dphi, d, B, (int)per_z, (long long)M * s.Hp * s.Wp * s.cout, s.Ho, s.Wo, s.cout,
                                               FastDiv((uint32_t)s.cout), FastDiv((uint32_t)s.Wo), FastDiv((uint32_t)s.Ho));
    LIP_LAUNCH_CHECK();
    return LIP_OK;
  }
  unpool_mask_kernel<<<ew_grid(total), 256, 0, st>>>(tin, s.dphi, d, total, per_z, s.Ho, s.Wo, s.cout, s.pool);
  LIP_LAUNCH_CHECK();
  return LIP_OK;
}

// ---- workspace --------------------------------------------------------------------------------------------------------
struct CnnWs {
  void* tail_ws = nullptr;   // workspace of the dense tail's sweeps (lip_model::tail)
  size_t tail_bytes = 0;
  float* t[2];    // stage outputs / tangent images, ping-pong          [B * max_s M*out_per_point]
  float* raw;     // pre-pool stage output / delta w.r.t. pre-activation [B * max_s M*P*cout]   (x2: delta ping-pong)
  float* raw2;
  float* col;     // im2col of a tangent image / delta . W^T             [B * max_{conv s>0} M*P*Kc]
};

struct CnnSizes { size_t t, raw, col; };

CnnSizes cnn_sizes(const lip_model* m, int64_t B) {
  CnnSizes z{0, 0, 0};
  const int nS = (int)m->CS.size();
  for (int i = 0; i < nS; ++i) {
    const ConvStage& s = m->CS[i];
    const size_t t = (size_t)B * m->M * (size_t)s.out_per_point();
    const size_t raw = (size_t)B * m->M * (size_t)s.P * s.cout;
    // the gradient w.r.t. a stage's INPUT image is also staged in t[]
    const size_t tin = (size_t)B * m->M * (size_t)(s.type == 1 ? (size_t)s.Hi * s.Wi * s.cin : (size_t)s.Kc);
    z.t = t > z.t ? t : z.t;
    z.t = (i > 0 && tin > z.t) ? tin : z.t;
    z.raw = raw > z.raw ? raw : z.raw;
    if (i > 0) {
      const size_t col = (size_t)B * m->M * (size_t)s.P * s.Kc;
      z.col = col > z.col ? col : z.col;
    }
  }
  z.t = align_up(z.t, 64); z.raw = align_up(z.raw, 64); z.col = align_up(z.col, 64);
  return z;
}

int cnn_carve(const lip_model* m, int64_t B, void* ws, size_t bytes, CnnWs* w) {
  const size_t need = cnn_ws_bytes(m, B);
  if (bytes < need || ws == nullptr) {
    set_error("workspace too small: need %zu bytes, got %zu", need, bytes);
    return LIP_ERR_WORKSPACE;
  }
  const CnnSizes z = cnn_sizes(m, B);
  float* base = (float*)align_up((uintptr_t)ws, 256);
  w->t[0] = base; w->t[1] = base + z.t;
  w->raw = base + 2 * z.t; w->raw2 = w->raw + z.raw;
  w->col = w->raw2 + z.raw;
  if (m->tail_on) {
    w->tail_ws = (void*)align_up((uintptr_t)(w->col + z.col), 256);
    w->tail_bytes = mlp_tail_ws_bytes(m->tail, B);
  }
  return LIP_OK;
}

// ---- JVP sweep: V[B, D] -> dlogits [B, M, K] in dst ------------------------------------------------------------------
int cnn_jvp_sweep(lip_model* m, const float* V, int64_t B, const CnnWs& w, float* dst, cudaStream_t st) {
  const int nS = (int)m->CS.size();
  const float* T = nullptr;   // tangent of the stage input, [B, M, in_per_point]
  for (int i = 0; i < nS; ++i) {
    const ConvStage& s = m->CS[i];
    const bool last = (i == nS - 1);
    const int64_t R = m->M * (int64_t)s.P;
    if (m->tail_on && i == m->tail_first)       // the dense tail on the MLP sweeps (tcgen05 for the wide layers)
      return mlp_tail_jvp(m->tail, V, m->D, T, B, w.tail_ws, w.tail_bytes, dst, st);
    if (!last && cnn_stage_fusable(m, i)) {     // conv + mask + pool in one kernel, no patch buffer
      int rc = cnn_fused_jvp(m, i, V, m->D, T, w.t[i & 1], B, st);
      if (rc) return rc;
      T = w.t[i & 1];
      continue;
    }
    const float* A2 = nullptr;
    if (i > 0) {
      if (s.type == 1) {
        int rc = launch_im2col(T, w.col, B * m->M, s, st);
        if (rc) return rc;
        A2 = w.col;
      } else {
        A2 = T;
      }
    }
    float* out = last ? dst : (s.pool ? w.raw : w.t[i & 1]);
    GemmProblem p;
    p.M = R; p.N = s.cout; p.K = s.Kc; p.batch = B;
    p.A1 = {s.Aop, 0, s.Kc, 1};
    p.B1 = {V + s.woff, m->D, s.cout, 1};
    if (A2) {
      p.A2 = {A2, R * (int64_t)s.Kc, s.Kc, 1};
      p.B2 = {m->theta + s.woff, 0, s.cout, 1};
      p.K2 = s.Kc;
    }
    p.C = out; p.c_sz = R * (int64_t)s.cout; p.c_sm = s.cout;
    p.epi.bias = V + s.boff; p.epi.bias_sz = m->D;
    if (!last) { p.epi.mask = s.dphi; p.epi.mask_sm = s.cout; }
    int rc = gemm_simt(p, st);
    if (rc) return rc;
    if (!last && s.pool) {
      rc = launch_avgpool(w.raw, w.t[i & 1], B * m->M, s, st);
      if (rc) return rc;
    }
    T = last ? nullptr : w.t[i & 1];
  }
  return LIP_OK;
}

// ---- VJP sweep: delta at the logits ([B, M, K] contiguous, in `dl`) -> out[B, D] = scale * J^T delta + add_scale * add ------
int cnn_vjp_sweep(lip_model* m, const float* dl, int64_t B, const CnnWs& w, float* out, float scale, const float* add,
                  float add_scale, cudaStream_t st) {
  const int nS = (int)m->CS.size();
  const size_t col_elems = cnn_sizes(m, B).col;
  const float* d = dl;     // delta w.r.t. the pre-activation of stage i, [B, M*P, cout]
  const float* tin_f = nullptr;   // set when stage i runs fused: the gradient w.r.t. its pooled output, [B, M, out_per_point]
  int istart = nS - 1;
  if (m->tail_on) {               // the dense tail on the MLP sweeps; it hands back the cotangent of its input features
    float* cot = m->tail_first > 0 ? w.t[0] : nullptr;
    int rc = mlp_tail_vjp(m->tail, dl, B, w.tail_ws, w.tail_bytes, out, m->D, scale, add, m->D, add_scale, cot, st);
    if (rc) return rc;
    if (m->tail_first == 0) return LIP_OK;
    istart = m->tail_first - 1;
    if (cnn_stage_fusable(m, istart)) {
      tin_f = cot;
    } else {
      rc = launch_unpool_mask(cot, w.raw, B, m->M, m->CS[istart], st);
      if (rc) return rc;
      d = w.raw;
    }
  }
  for (int i = istart; i >= 0; --i) {
    const ConvStage& s = m->CS[i];
    const int64_t R = m->M * (int64_t)s.P;
    const float* tin = nullptr;     // gradient w.r.t. the output of stage i-1, [B, M, out_per_point(i-1)]
    if (tin_f) {                    // unpool + mask, kernel / bias gradients and the delta back-propagation in one kernel
      float* gin = (i > 0) ? (tin_f == w.t[0] ? w.t[1] : w.t[0]) : nullptr;
      int rc = cnn_fused_vjp(m, i, tin_f, gin, out, B, scale, add, add_scale, w.col, (int64_t)col_elems, st);
      if (rc) return rc;
      if (i == 0) break;
      tin = gin;
    } else {
      {  // weight gradient [Kc x cout] = Aop^T [Kc x R] . d [R x cout], written in place into out[b, woff ...]
        GemmProblem p;
        p.M = s.Kc; p.N = s.cout; p.K = R; p.batch = B;
        p.A1 = {s.Aop, 0, 1, s.Kc};
        p.B1 = {d, R * (int64_t)s.cout, s.cout, 1};
        p.C = out + s.woff; p.c_sz = m->D; p.c_sm = s.cout;
        p.epi.scale = scale;
        if (add) { p.epi.add = add + s.woff; p.epi.add_sz = m->D; p.epi.add_scale = add_scale; }
        p.splitk_ws = w.col; p.splitk_ws_elems = (int64_t)col_elems;      // col is free here (used again by the G GEMM below)
        int rc = gemm_simt(p, st);
        if (rc) return rc;
        rc = launch_bias_grad(d, nullptr, R, s.cout, s.cout, B, out + s.boff, m->D, scale, add ? add + s.boff : nullptr, m->D,
                              add_scale, st);
        if (rc) return rc;
      }
      if (i == 0) break;
      const ConvStage& sp = m->CS[i - 1];
      // G = d . W^T : [R x Kc] per probe
      const bool fuse_mask = (s.type == 0 && !sp.pool && sp.P == 1);      // dense after dense: mask in the epilogue
      float* G = (s.type == 1) ? w.col : (fuse_mask ? ((d == w.raw) ? w.raw2 : w.raw) : w.t[0]);
      {
        GemmProblem p;
        p.M = R; p.N = s.Kc; p.K = s.cout; p.batch = B;
        p.A1 = {d, R * (int64_t)s.cout, s.cout, 1};
        p.B1 = {m->theta + s.woff, 0, 1, s.cout};
        p.C = G; p.c_sz = R * (int64_t)s.Kc; p.c_sm = s.Kc;
        if (fuse_mask) { p.epi.mask = sp.dphi; p.epi.mask_sm = s.Kc; }
        int rc = gemm_simt(p, st);
        if (rc) return rc;
      }
      if (fuse_mask) { d = G; continue; }
      tin = G;
      if (s.type == 1) {
        int rc = launch_col2im(w.col, w.t[0], B * m->M, s, st);
        if (rc) return rc;
        tin = w.t[0];
      }
    }
    if (cnn_stage_fusable(m, i - 1)) { tin_f = tin; continue; }
    tin_f = nullptr;
    float* dn = (d == w.raw) ? w.raw2 : w.raw;
    int rc = launch_unpool_mask(tin, dn, B, m->M, m->CS[i - 1], st);
    if (rc) return rc;
    d = dn;
  }
  return LIP_OK;
}

}  // namespace

namespace lip {

int cnn_parse(lip_model* m, const lip_layer_desc* layers, int32_t n_layers, int64_t num_params) {
  const lip_layer_desc& in = layers[0];
  LIP_REQUIRE(in.kh > 0 && in.kw > 0 && in.in_features > 0, "lip_model_create: INPUT needs height, width, channels > 0");
  m->is_cnn = true;
  m->in_h = in.kh; m->in_w = in.kw; m->in_c = in.in_features;
  int H = in.kh, W = in.kw, C = in.in_features;
  bool image = true;
  int64_t flat = 0;
  int pend_pad = 0;
  int64_t counted = 0;
  for (int i = 1; i < n_layers; ++i) {
    const lip_layer_desc& d = layers[i];
    switch (d.op) {
      case LIP_OP_ZEROPAD:
        LIP_REQUIRE(image && d.pad >= 0, "lip_model_create: ZEROPAD at %d needs an image input", i);
        pend_pad += d.pad;
        break;
      case LIP_OP_CONV2D: {
        LIP_REQUIRE(image, "lip_model_create: CONV2D at %d after FLATTEN", i);
        LIP_REQUIRE(d.stride == 1, "lip_model_create: CONV2D at %d: only stride 1 is built (got %d)", i, d.stride);
        LIP_REQUIRE(d.in_features == C && d.out_features > 0 && d.kh > 0 && d.kw > 0,
                    "lip_model_create: CONV2D at %d: cin %d != current channels %d (or bad shape)", i, d.in_features, C);
        const int pad = pend_pad + d.pad;
        const int Ho = H + 2 * pad - d.kh + 1, Wo = W + 2 * pad - d.kw + 1;
        LIP_REQUIRE(Ho > 0 && Wo > 0, "lip_model_create: CONV2D at %d: kernel larger than the padded image", i);
        const int64_t nk = (int64_t)d.kh * d.kw * C * d.out_features;
        LIP_REQUIRE(d.bias_offset >= 0 && d.kernel_offset >= 0 && d.bias_offset + d.out_features <= num_params &&
                        d.kernel_offset + nk <= num_params,
                    "lip_model_create: CONV2D at %d has an invalid offset (flax nn.Conv with use_bias=True expected)", i);
        LIP_REQUIRE(m->CS.empty() || m->CS.back().act >= 0 || m->CS.back().pool, "lip_model_create: two affine ops in a row at %d", i);
        ConvStage s;
        s.type = 1; s.Hi = H; s.Wi = W; s.cin = C; s.pad = pad; s.kh = d.kh; s.kw = d.kw; s.Ho = Ho; s.Wo = Wo;
        s.cout = d.out_features; s.Hp = Ho; s.Wp = Wo; s.boff = d.bias_offset; s.woff = d.kernel_offset;
        s.P = Ho * Wo; s.Kc = d.kh * d.kw * C;
        m->CS.push_back(s);
        counted += nk + d.out_features;
        H = Ho; W = Wo; C = d.out_features; pend_pad = 0;
        break;
      }
      case LIP_OP_TANH: case LIP_OP_GELU_TANH: case LIP_OP_RELU:
        LIP_REQUIRE(!m->CS.empty() && m->CS.back().act == -1 && !m->CS.back().pool,
                    "lip_model_create: activation at %d must directly follow CONV2D / DENSE", i);
        m->CS.back().act = d.op;
        break;
      case LIP_OP_AVGPOOL2: {
        LIP_REQUIRE(image && !m->CS.empty() && m->CS.back().type == 1 && !m->CS.back().pool && pend_pad == 0,
                    "lip_model_create: AVGPOOL2 at %d must follow a CONV2D stage", i);
        ConvStage& s = m->CS.back();
        LIP_REQUIRE(s.Ho % 2 == 0 && s.Wo % 2 == 0, "lip_model_create: AVGPOOL2 at %d on an odd %dx%d image", i, s.Ho, s.Wo);
        s.pool = 1; s.Hp = s.Ho / 2; s.Wp = s.Wo / 2;
        H = s.Hp; W = s.Wp;
        break;
      }
      case LIP_OP_FLATTEN:
        LIP_REQUIRE(image && pend_pad == 0, "lip_model_create: FLATTEN at %d needs an image", i);
        image = false; flat = (int64_t)H * W * C;
        break;
      case LIP_OP_DENSE: {
        if (image) { LIP_REQUIRE(pend_pad == 0, "lip_model_create: ZEROPAD before DENSE at %d", i); image = false; flat = (int64_t)H * W * C; }
        LIP_REQUIRE(d.in_features == flat && d.out_features > 0, "lip_model_create: DENSE at %d: in_features %d != %lld", i,
                    d.in_features, (long long)flat);
        LIP_REQUIRE(d.bias_offset >= 0 && d.kernel_offset >= 0 && d.bias_offset + d.out_features <= num_params &&
                        d.kernel_offset + (int64_t)d.in_features * d.out_features <= num_params,
                    "lip_model_create: DENSE at %d has an invalid offset", i);
        LIP_REQUIRE(m->CS.empty() || m->CS.back().act >= 0 || m->CS.back().pool, "lip_model_create: two affine ops in a row at %d", i);
        ConvStage s;
        s.type = 0; s.cin = d.in_features; s.cout = d.out_features; s.Kc = d.in_features; s.P = 1;
        s.boff = d.bias_offset; s.woff = d.kernel_offset;
        m->CS.push_back(s);
        counted += (int64_t)d.in_features * d.out_features + d.out_features;
        flat = d.out_features;
        break;
      }
      default:
        set_error("lip_model_create: unsupported op %d at position %d of a conv program", d.op, i);
        return LIP_ERR_INVALID;
    }
  }
  LIP_REQUIRE(!m->CS.empty() && m->CS.back().type == 0 && m->CS.back().act == -1 && !image,
              "lip_model_create: a conv program must end with a DENSE layer");
  for (size_t i = 0; i + 1 < m->CS.size(); ++i)
    LIP_REQUIRE(m->CS[i].act >= 0, "lip_model_create: stage %zu has no activation", i);
  LIP_REQUIRE(counted == num_params, "lip_model_create: layers hold %lld parameters but num_params = %lld", (long long)counted,
              (long long)num_params);
  m->K = m->CS.back().cout;
  {  // trailing dense stages as a dense program of their own (see lip_model::tail)
    int i0 = (int)m->CS.size();
    while (i0 > 0 && m->CS[i0 - 1].type == 0) --i0;
    if (m->tail) { lip_model_destroy(m->tail); m->tail = nullptr; }
    m->tail_first = i0;
    if (i0 < (int)m->CS.size()) {
      m->tail = mlp_make_tail(m->CS, i0, m->model_type, m->D);
      LIP_REQUIRE(m->tail != nullptr, "lip_model_create: out of host memory");
    }
  }
  return LIP_OK;
}

int cnn_bind(lip_model* m, const float* theta, const float* Z, int64_t M, cudaStream_t st) {
  const int nS = (int)m->CS.size();
  size_t max_raw = 0, max_x = (size_t)M * m->in_h * m->in_w * m->in_c;
  for (int i = 0; i < nS; ++i) {
    ConvStage& s = m->CS[i];
    const size_t R = (size_t)M * s.P;
    LIP_CHECK_CUDA(cudaMalloc(&s.Aop, sizeof(float) * R * s.Kc + 256));
    if (i < nS - 1) LIP_CHECK_CUDA(cudaMalloc(&s.dphi, sizeof(float) * R * s.cout + 256));
    max_raw = R * s.cout > max_raw ? R * s.cout : max_raw;
    const size_t x = (size_t)M * (size_t)s.out_per_point();
    max_x = x > max_x ? x : max_x;
  }
  LIP_CHECK_CUDA(cudaMalloc(&m->cnn_tmp_out, sizeof(float) * max_raw + 256));
  LIP_CHECK_CUDA(cudaMalloc(&m->cnn_tmp_x, sizeof(float) * max_x + 256));
  LIP_CHECK_CUDA(cudaMalloc(&m->logits, sizeof(float) * (size_t)M * m->K));
  LIP_CHECK_CUDA(cudaMalloc(&m->P, sizeof(float) * (size_t)M * m->K));
  LIP_CHECK_CUDA(cudaMalloc(&m->S, sizeof(float) * (size_t)M * m->K));
  const float* X = Z;       // current stage input, [M, in_per_point]
  for (int i = 0; i < nS; ++i) {
    ConvStage& s = m->CS[i];
    const bool last = (i == nS - 1);
    const int64_t R = M * (int64_t)s.P;
    if (s.type == 1) {
      int rc = launch_im2col(X, s.Aop, M, s, st);
      if (rc) return rc;
      const size_t xin = sizeof(float) * (size_t)M * s.Hi * s.Wi * s.cin;
      LIP_CHECK_CUDA(cudaMalloc(&s.Xin, xin + 256));
      LIP_CHECK_CUDA(cudaMemcpyAsync(s.Xin, X, xin, cudaMemcpyDeviceToDevice, st));
    } else if (X != s.Aop) {
      LIP_CHECK_CUDA(cudaMemcpyAsync(s.Aop, X, sizeof(float) * (size_t)R * s.Kc, cudaMemcpyDeviceToDevice, st));
    }
    // where the stage output goes: straight into the next dense stage's operand, else a temp image
    float* xn = last ? m->logits : (m->CS[i + 1].type == 0 ? m->CS[i + 1].Aop : m->cnn_tmp_x);
    float* out = s.pool ? m->cnn_tmp_out : xn;
    GemmProblem p;
    p.M = R; p.N = s.cout; p.K = s.Kc; p.batch = 1;
    p.A1 = {s.Aop, 0, s.Kc, 1};
    p.B1 = {theta + s.woff, 0, s.cout, 1};
    p.C = out; p.c_sz = 0; p.c_sm = s.cout;
    p.epi.bias = theta + s.boff; p.epi.bias_sz = 0;
    if (!last) { p.epi.act = s.act; p.epi.dphi_out = s.dphi; }
    int rc = gemm_simt(p, st);
    if (rc) return rc;
    if (s.pool) {
      rc = launch_avgpool(out, xn, M, s, st);
      if (rc) return rc;
    }
    X = xn;
  }
  if (m->model_type == LIP_CLASSIFIER) {
    int rc = launch_softmax(m->logits, m->P, m->S, M, m->K, st);
    if (rc) return rc;
  }
  m->tc_on = false;
  m->tail_on = false;
  static const bool tail_off = getenv("LIP_CNN_TC_TAIL") && atoi(getenv("LIP_CNN_TC_TAIL")) == 0;
  if (m->tail && !tail_off && m->use_tc != 0) {
    m->tail->use_tc = m->use_tc;
    int rc = lip_model_bind(m->tail, theta, m->CS[m->tail_first].Aop, M, m->logvar, (lip_stream_t)st);
    if (rc) return rc;
    m->tail_on = m->tail->tc_on;       // worth it only when some layer reaches the tensor cores; else the stage path's skinny kernels
  }
  m->bound = true;
  return LIP_OK;
}

size_t cnn_ws_bytes(const lip_model* m, int64_t B) {
  const CnnSizes z = cnn_sizes(m, B);
  return (2 * z.t + 2 * z.raw + z.col) * sizeof(float) + 512 + (m->tail_on ? align_up(mlp_tail_ws_bytes(m->tail, B), 256) + 512 : 0);
}

int cnn_ggn_vp(lip_model* m, const float* V, float* out, int64_t B, float recal, float alpha, void* ws, size_t bytes,
               cudaStream_t st) {
  CnnWs w;
  int rc = cnn_carve(m, B, ws, bytes, &w);
  if (rc) return rc;
  float* dl = w.raw2;                       // [B, M, K]: not touched by the JVP sweep
  rc = cnn_jvp_sweep(m, V, B, w, dl, st);
  if (rc) return rc;
  if (m->model_type == LIP_CLASSIFIER) {
    rc = launch_factor(dl, dl, m, B, 0, 1.f, st);
    if (rc) return rc;
  }
  return cnn_vjp_sweep(m, dl, B, w, out, recal, alpha != 0.f ? V : nullptr, alpha, st);
}

int cnn_wt_apply(lip_model* m, const float* V, float* out, int64_t B, float scale, int32_t factor, void* ws, size_t bytes,
                 cudaStream_t st) {
  CnnWs w;
  int rc = cnn_carve(m, B, ws, bytes, &w);
  if (rc) return rc;
  rc = cnn_jvp_sweep(m, V, B, w, out, st);
  if (rc) return rc;
  float s = scale;
  if (factor == LIP_FACTOR_SQRT && m->model_type == LIP_REGRESSOR) s *= expf(-0.5f * m->logvar);
  if (factor == LIP_FACTOR_SQRT && m->model_type == LIP_CLASSIFIER) return launch_factor(out, out, m, B, 1, s, st);
  if (s != 1.f) return launch_scale_copy(out, out, B * m->M * m->K, s, st);
  return LIP_OK;
}

// ---- gradients with respect to the input images Z (lip_zgrad for conv stage programs) ---------------------------
// Recurrences of lip_zgrad.cu with phi = relu, so phi'' = 0 and no tangent has to be kept:
//     e_{i-1} = phi'_{i-1} * pool^T( col2im( e_i W_i^T ) )
//     q_{i-1} = phi'_{i-1} * pool^T( col2im( e_i dW_i[b]^T + q_i W_i^T ) )
//     dZ      = col2im_0( sum_probes  e_0 dW_0[b]^T + q_0 W_0^T )
// i.e. two delta back-propagations (the second with a dual-K GEMM that also reads the probe's own kernel block in place),
// continued through stage 0 down to the input image.
struct CnnZSizes { size_t raw, fin, sum, small, total; };

static CnnZSizes cnn_zsizes(const lip_model* m, int nseg, int64_t B, int per_probe) {
  CnnZSizes z;
  const CnnSizes c = cnn_sizes(m, B);
  const ConvStage& s0 = m->CS[0];
  z.raw = c.raw;
  z.fin = align_up((size_t)nseg * B * m->M * (size_t)s0.P * s0.Kc, 64);
  z.sum = align_up((size_t)(per_probe ? B : 1) * m->M * (size_t)s0.P * s0.Kc, 64);
  z.small = align_up((size_t)nseg * B * m->M * m->K, 64);
  z.total = align_up(cnn_ws_bytes(m, B), 256) + (2 * z.raw + z.fin + z.sum + 3 * z.small) * sizeof(float) + 512;
  return z;
}

size_t cnn_zgrad_ws_bytes(const lip_model* m, int32_t mode, int64_t B) {
  return cnn_zsizes(m, mode == LIP_ZGRAD_GGN ? 2 : 1, B, 1).total;
}

int cnn_zgrad(lip_model* m, int32_t mode, const float* X1, const float* X2, float* out, int64_t B, float scale, int32_t per_probe,
              void* ws, size_t bytes, cudaStream_t st) {
  const int nS = (int)m->CS.size();
  for (int i = 0; i + 1 < nS; ++i)
    if (m->CS[i].act != LIP_OP_RELU) {
      set_error("lip_zgrad: conv stage programs need relu activations (stage %d has op %d)", i, m->CS[i].act);
      return LIP_ERR_UNSUPPORTED;
    }
  const int nseg = mode == LIP_ZGRAD_GGN ? 2 : 1;
  const CnnZSizes z = cnn_zsizes(m, nseg, B, 1);
  if (!ws || bytes < z.total) {
    set_error("lip_zgrad: workspace too small: need %zu bytes, got %zu", z.total, bytes);
    return LIP_ERR_WORKSPACE;
  }
  CnnWs w;
  const size_t inner = align_up(cnn_ws_bytes(m, B), 256);
  int rc = cnn_carve(m, B, ws, inner, &w);
  if (rc) return rc;
  float* base = (float*)align_up((uintptr_t)ws + inner, 256);
  float* qbuf[2] = {base, base + z.raw};
  float* fin = base + 2 * z.raw;
  float* sum = fin + z.fin;
  float* dl = sum + z.sum;
  float* Cc = dl + z.small;
  float* Gf = Cc + z.small;
  float* ebuf[2] = {w.raw, w.raw2};
  const float* Vseg[2] = {X1, mode == LIP_ZGRAD_GGN ? X2 : nullptr};
  const int64_t M = m->M;
  const int64_t small_seg = B * M * m->K;

  for (int sg = 0; sg < nseg; ++sg) {   // forward tangent pass: only the tangent logits are needed (relu: phi'' = 0)
    rc = cnn_jvp_sweep(m, Vseg[sg], B, w, dl + sg * small_seg, st);
    if (rc) return rc;
  }
  rc = launch_zgrad_rows(mode, m, dl, X2, Cc, Gf, B, scale, st);
  if (rc) return rc;

  const ConvStage& s0 = m->CS[0];
  const int64_t fin_seg = B * M * (int64_t)s0.P * s0.Kc;
  for (int sg = 0; sg < nseg; ++sg) {
    const float* e = Cc + sg * small_seg;
    const float* q = Gf + sg * small_seg;
    int pp = 0;
    for (int i = nS - 1; i >= 0; --i) {
      const ConvStage& s = m->CS[i];
      const int64_t R = M * (int64_t)s.P;
      for (int which = 0; which < 2; ++which) {      // 0: the e path (shared kernel), 1: the q path (dual-K with the probe's kernel)
        if (i == 0 && which == 0) continue;          // stage 0 needs only  e_0 dW_0^T + q_0 W_0^T
        float* G = (i == 0) ? fin + sg * fin_seg : ((s.type == 1) ? w.col : w.t[0]);
        GemmProblem p;
        p.M = R; p.N = s.Kc; p.K = s.cout; p.batch = B;
        p.A1 = {e, R * (int64_t)s.cout, s.cout, 1};
        if (which == 0) {
          p.B1 = {m->theta + s.woff, 0, 1, s.cout};
        } else {
          p.B1 = {Vseg[sg] + s.woff, m->D, 1, s.cout};
          p.A2 = {q, R * (int64_t)s.cout, s.cout, 1};
          p.B2 = {m->theta + s.woff, 0, 1, s.cout};
          p.K2 = s.cout;
        }
        p.C = G; p.c_sz = R * (int64_t)s.Kc; p.c_sm = s.Kc;
        rc = gemm_simt(p, st);
        if (rc) return rc;
        if (i == 0) break;
        const ConvStage& sp = m->CS[i - 1];
        const float* tin = G;     // gradient w.r.t. the output of stage i-1, [B, M, out_per_point(i-1)]
        if (s.type == 1) {
          rc = launch_col2im(w.col, w.t[0], B * M, s, st);
          if (rc) return rc;
          tin = w.t[0];
        }
        rc = launch_unpool_mask(tin, which == 0 ? ebuf[pp] : qbuf[pp], B, M, sp, st);
        if (rc) return rc;
      }
      if (i > 0) { e = ebuf[pp]; q = qbuf[pp]; pp ^= 1; }
    }
  }
  // sum over probes (per probe in GGN mode: the two halves), then back to the image through stage 0's patches
  const int64_t per0 = M * (int64_t)s0.P * s0.Kc;
  const float* src = fin;
  int64_t images = M;
  if (!per_probe) {
    rc = launch_batch_sum(fin, sum, per0, (int64_t)nseg * B, st);
    if (rc) return rc;
    src = sum;
  } else {
    images = B * M;
    if (nseg == 2) {
      rc = launch_batch_sum(fin, sum, B * per0, 2, st);
      if (rc) return rc;
      src = sum;
    }
  }
  if (s0.type == 1) return launch_col2im(src, out, images, s0, st);
  return launch_scale_copy(src, out, images * (int64_t)s0.Kc, 1.f, st);
}

int cnn_w_apply(lip_model* m, const float* U, float* out, int64_t B, float scale, int32_t factor, const float* add,
                float add_scale, void* ws, size_t bytes, cudaStream_t st) {
  CnnWs w;
  int rc = cnn_carve(m, B, ws, bytes, &w);
  if (rc) return rc;
  float s = scale;
  float* dl = w.raw2;
  if (factor == LIP_FACTOR_SQRT && m->model_type == LIP_CLASSIFIER) {
    rc = launch_factor(U, dl, m, B, 2, 1.f, st);
  } else {
    if (factor == LIP_FACTOR_SQRT) s *= expf(-0.5f * m->logvar);
    rc = launch_scale_copy(U, dl, B * m->M * m->K, 1.f, st);
  }
  if (rc) return rc;
  return cnn_vjp_sweep(m, dl, B, w, out, s, add, add_scale, st);
}

}  // namespace lip
