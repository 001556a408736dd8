// lip_zgrad.cu — gradients of the hot-path operators with respect to the inducing points Z.
//
// The reference trains Z with jax.value_and_grad of the KL objective (src/train_inducing.py:195-232), i.e. JAX
// differentiates THROUGH ggn_vp / Wfun / WTfun (src/ggn.py:9-146) with respect to Z.  A custom call has no autodiff rule,
// so these are the VJP-with-respect-to-Z rules a jax.custom_vjp around lip_ggn_vp / lip_w_apply / lip_wt_apply needs
// (the rule with respect to the vector argument is the operator itself: GGN is symmetric, W and W^T are adjoint).
//
// With f_i = f(z_i; theta) the model outputs, J_i = df_i/dtheta and a probe's tangent v = (dW_l, db_l):
//   forward tangent pass (stores dh_l):    dh_l = A_l dW_l + dA_l W_l + db_l,   dA_{l+1} = phi'(h_l) * dh_l,   (J_i v) = dh_L
//   scalar being differentiated:           s = sum_i  c_i . (J_i v)  +  g_i . f_i            (c, g constants, built per mode)
//   reverse pass (e_l = ds/d dh_l, q_l = ds/d h_l; e_L = c, q_L = g):
//        X       = e_l W_l^T                                             (adjoint of dA_l)
//        e_{l-1} = phi'_{l-1} * X
//        q_{l-1} = phi'_{l-1} * (e_l dW_l^T + q_l W_l^T) + phi''_{l-1} * dh_{l-1} * X
//        dZ      = sum_probes (e_0 dW_0^T + q_0 W_0^T)
// Every contraction is one strided batched GEMM over all (probe, point) pairs (gemm_simt, fp32 FMA), the per-probe weight
// block dW_l[b] is read in place from the caller's [B, D] array exactly as in the JVP sweep (lip_model.cu).
// The output-space parts (softmax Hessian H = diag(p) - pp^T, its factor L and their derivatives with respect to the
// logits) are K-wide row kernels in registers.
//
//   LIP_ZGRAD_GGN   s = ubar^T (scale * sum_i J_i^T H_i J_i v)                    X1 = ubar [B,D], X2 = v [B,D]
//   LIP_ZGRAD_WT    s = sum_i Ybar_i . (scale * L_i^T J_i v)                       X1 = v [B,D],   X2 = Ybar [B,M,K]
//   LIP_ZGRAD_W     s = ubar^T (scale * sum_i J_i^T L_i U_i)                       X1 = ubar [B,D], X2 = U [B,M,K]
//   LIP_ZGRAD_JVP   s = sum_i C_i . (scale * J_i v)   (factor NONE, lla.py:153)    X1 = v [B,D],   X2 = C [B,M,K]
// Dense programs (models M1 / M2) here; relu conv stage programs (LeNet5, M3) in lip_cnn.cu (cnn_zgrad); residual programs return
// LIP_ERR_UNSUPPORTED.
//
// Two executions of the same recurrences: zgrad_simt (fp32 FMA GEMMs, any activation, any width) and zgrad_tc, which runs the
// layers the model already has on the tcgen05 path (lip_model.cu: in, out >= 64) through the 3xTF32 tensor-core GEMMs
// (lip_gemm_tc.cu): the forward tangent pass IS the JVP sweep (keeping every layer's masked tangent dA as a TF32 (hi, lo) pair),
// X = e W^T is the delta-backprop instance, and q_{l-1} is ONE dual-K GEMM [e_l | q_l] x [dW_l^T ; W_l^T] whose epilogue applies
// the phi' mask and adds T = (phi''/phi') * dA * X (for tanh phi''/phi' = -2 tanh(h); relu: 0 - zgrad_tc is used for those two).
#include <math.h>
#include <stdlib.h>

#include "lip_model.cuh"

namespace lip {
namespace {

__device__ __forceinline__ float act_second(int act, float h) {
  if (act == LIP_OP_TANH) {
    const float a = tanhf(h);
    return -2.f * a * (1.f - a * a);
  }
  if (act == LIP_OP_RELU) return 0.f;
  const float c = 0.7978845608028654f, k = 0.044715f;   // gelu, tanh approximation
  const float u = c * (h + k * h * h * h), t = tanhf(u);
  const float up = c * (1.f + 3.f * k * h * h), upp = 6.f * c * k * h;
  const float sech2 = 1.f - t * t;
  return sech2 * up + 0.5f * h * sech2 * (upp - 2.f * t * up * up);
}

// h (pre-activations, [n]) -> phi''(h) in place
__global__ void act_second_kernel(float* __restrict__ h, int64_t n, int act) {
  int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i < n) h[i] = act_second(act, h[i]);
}

// out[z][i] = mask[i] * x[z][i]      (i < per, mask shared by all z)
__global__ void mask_mul_kernel(const float* __restrict__ x, const float* __restrict__ mask, float* __restrict__ out,
                                int64_t per, int64_t total) {
  int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i < total) out[i] = x[i] * __ldg(mask + i % per);
}

// e[z][i] = dphi[i] * X[z][i];   T[z][i] = ddphi[i] * dh[z][i] * X[z][i]   (X may alias e)
__global__ void reverse_act_kernel(const float* __restrict__ X, const float* __restrict__ dh, const float* __restrict__ dphi,
                                   const float* __restrict__ ddphi, float* __restrict__ e, float* __restrict__ T, int64_t per,
                                   int64_t total) {
  int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i >= total) return;
  const int64_t j = i % per;
  const float x = X[i];
  T[i] = __ldg(ddphi + j) * dh[i] * x;
  e[i] = __ldg(dphi + j) * x;
}

// out[i] = scale * sum_z x[z][i]   (fixed order: deterministic)
__global__ void batch_sum_kernel(const float* __restrict__ x, float* __restrict__ out, int64_t per, int64_t nb, float scale) {
  int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i >= per) return;
  float acc = 0.f;
  for (int64_t z = 0; z < nb; ++z) acc += x[z * per + i];
  out[i] = scale * acc;
}

// Output-space rows: from the tangent logits dl [nb, M, K] (and the caller's X2 [B, M, K]) build e_L = c and q_L = g.
//   mode GGN: rows z < B carry a = J ubar, rows z >= B carry b = J v (nb = 2B)
__global__ void zgrad_rows_kernel(int mode, int classifier, const float* __restrict__ dl, const float* __restrict__ X2,
                                  const float* __restrict__ P, const float* __restrict__ S, float* __restrict__ Cc,
                                  float* __restrict__ Gf, int64_t B, int64_t M, int K, float scale) {
  const int64_t rows = B * M;
  int64_t r = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (r >= rows) return;
  const int64_t i = r % M;
  const float* p = P + i * K;
  const float* s = S + i * K;
  if (mode == LIP_ZGRAD_JVP || !classifier) {
    // no output-space dependence on f: regressor H / L are constants (folded into `scale` by the caller of this kernel)
    if (mode == LIP_ZGRAD_GGN) {
      const float* a = dl + r * K;
      const float* b = dl + (rows + r) * K;
      for (int k = 0; k < K; ++k) {
        Cc[r * K + k] = scale * b[k];
        Cc[(rows + r) * K + k] = scale * a[k];
        Gf[r * K + k] = 0.f;
        Gf[(rows + r) * K + k] = 0.f;
      }
    } else {
      for (int k = 0; k < K; ++k) {
        Cc[r * K + k] = scale * X2[r * K + k];
        Gf[r * K + k] = 0.f;
      }
    }
    return;
  }
  if (mode == LIP_ZGRAD_GGN) {
    const float* a = dl + r * K;
    const float* b = dl + (rows + r) * K;
    float pa = 0.f, pb = 0.f, pab = 0.f;
    for (int k = 0; k < K; ++k) { pa += p[k] * a[k]; pb += p[k] * b[k]; pab += p[k] * a[k] * b[k]; }
    // w = a*b - (p.b) a - (p.a) b ;  t = H w = p*w - p (p.w),   p.w = p.ab - 2 (p.a)(p.b)
    const float pw = pab - 2.f * pa * pb;
    for (int k = 0; k < K; ++k) {
      const float Hb = p[k] * (b[k] - pb), Ha = p[k] * (a[k] - pa);
      const float w = a[k] * b[k] - pb * a[k] - pa * b[k];
      Cc[r * K + k] = scale * Hb;
      Cc[(rows + r) * K + k] = scale * Ha;
      Gf[r * K + k] = 0.f;
      Gf[(rows + r) * K + k] = scale * p[k] * (w - pw);
    }
  } else if (mode == LIP_ZGRAD_WT) {
    // s = Ybar . L^T b,  L^T b = r*b - (p.b) r  (r = sqrt p);  c = L Ybar = r*y - (r.y) p
    const float* y = X2 + r * K;
    const float* b = dl + r * K;
    float ry = 0.f, pb = 0.f, ryb = 0.f;
    for (int k = 0; k < K; ++k) { ry += s[k] * y[k]; pb += p[k] * b[k]; ryb += s[k] * y[k] * b[k]; }
    for (int k = 0; k < K; ++k) {
      const float Hb = p[k] * (b[k] - pb);
      const float g = 0.5f * (s[k] * y[k] * b[k] - p[k] * ryb) - ry * Hb - pb * 0.5f * (s[k] * y[k] - p[k] * ry);
      Cc[r * K + k] = scale * (s[k] * y[k] - ry * p[k]);
      Gf[r * K + k] = scale * g;
    }
  } else {  // LIP_ZGRAD_W:  s = a . L u,  L u = r*u - (r.u) p;  c = L u
    const float* u = X2 + r * K;
    const float* a = dl + r * K;
    float ru = 0.f, pa = 0.f, rau = 0.f;
    for (int k = 0; k < K; ++k) { ru += s[k] * u[k]; pa += p[k] * a[k]; rau += s[k] * a[k] * u[k]; }
    for (int k = 0; k < K; ++k) {
      const float Ha = p[k] * (a[k] - pa);
      const float g = 0.5f * (s[k] * a[k] * u[k] - p[k] * rau) - pa * 0.5f * (s[k] * u[k] - p[k] * ru) - ru * Ha;
      Cc[r * K + k] = scale * (s[k] * u[k] - ru * p[k]);
      Gf[r * K + k] = scale * g;
    }
  }
}

struct Seg { const float* V; int64_t b0, nb; };   // probes [b0, b0 + nb) of the working batch read their tangents from V

struct ZWs {
  std::vector<float*> dh;      // [nb, M, out_l], l < L
  float* buf[6];               // e / q ping-pong, X, T   (each [nb, M, wmax])
  float* dl;                   // [nb, M, K]
  float* Cc;                   // [nb, M, K]
  float* Gf;                   // [nb, M, K]
};

int64_t zg_wmax(const lip_model* m) {
  int64_t w = m->L[0].in;
  for (auto& L : m->L) w = L.out > w ? L.out : w;
  return w;
}

size_t zg_bytes(const lip_model* m, int64_t nb) {
  const int nL = (int)m->L.size();
  size_t fl = 0;
  for (int l = 0; l + 1 < nL; ++l) fl += align_up((size_t)nb * m->M * m->L[l].out, 64);
  fl += 6 * align_up((size_t)nb * m->M * zg_wmax(m), 64);
  fl += 3 * align_up((size_t)nb * m->M * m->K, 64);
  return fl * sizeof(float) + 512;
}

void zg_carve(const lip_model* m, int64_t nb, void* ws, ZWs* w) {
  const int nL = (int)m->L.size();
  float* base = (float*)align_up((uintptr_t)ws, 256);
  w->dh.assign(nL, nullptr);
  for (int l = 0; l + 1 < nL; ++l) {
    w->dh[l] = base; base += align_up((size_t)nb * m->M * m->L[l].out, 64);
  }
  const size_t per = align_up((size_t)nb * m->M * zg_wmax(m), 64);
  for (int i = 0; i < 6; ++i) { w->buf[i] = base; base += per; }
  const size_t pk = align_up((size_t)nb * m->M * m->K, 64);
  w->dl = base; base += pk;
  w->Cc = base; base += pk;
  w->Gf = base;
}

inline unsigned blocks(int64_t n) { return (unsigned)ceil_div(n, 256); }

// rho = phi'' / phi' at the bound points (0 where phi' vanishes: tanh saturated in fp32 / relu off, where phi'' * dh is 0 too)
__global__ void rho_kernel(float* __restrict__ ddphi, const float* __restrict__ dphi, int64_t n) {
  int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i < n) {
    const float d = dphi[i];
    ddphi[i] = d != 0.f ? ddphi[i] / d : 0.f;
  }
}

// rows x cols block with leading dimension ld (X, dA, e, T) against [M, cols] point-wise factors:
//   e = phi' * X  (as a TF32 (hi, lo) pair when e_lo != nullptr),   T = rho * (dA_hi + dA_lo) * X
__global__ void reverse_act_tc_kernel(const float* __restrict__ X, const float* __restrict__ dA_hi, const float* __restrict__ dA_lo,
                                      const float* __restrict__ dphi, const float* __restrict__ rho, float* __restrict__ e_hi,
                                      float* __restrict__ e_lo, float* __restrict__ T, int64_t rows, int64_t M, int cols, int ld) {
  // blockIdx.y strides over rows (no per-element division), threadIdx.x over 4-column groups (ld and the buffers are 16-byte aligned)
  const int c4 = (blockIdx.x * blockDim.x + threadIdx.x) * 4;
  if (c4 >= cols) return;
  const bool full = c4 + 4 <= cols;
  for (int64_t r = blockIdx.y; r < rows; r += gridDim.y) {
    const int64_t o = r * ld + c4, pm = (r % M) * cols + c4;
    float x[4], da[4], ph[4], rh[4];
    if (full) {
      *reinterpret_cast<float4*>(x) = *reinterpret_cast<const float4*>(X + o);
      *reinterpret_cast<float4*>(da) = *reinterpret_cast<const float4*>(dA_hi + o);
      if (dA_lo) {
        const float4 l = *reinterpret_cast<const float4*>(dA_lo + o);
        da[0] += l.x; da[1] += l.y; da[2] += l.z; da[3] += l.w;
      }
    } else {
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const bool ok = c4 + j < cols;
        x[j] = ok ? X[o + j] : 0.f;
        da[j] = ok ? dA_hi[o + j] + (dA_lo ? dA_lo[o + j] : 0.f) : 0.f;
      }
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) {            // [M, cols] factors: cols need not be a multiple of 4
      const bool ok = c4 + j < cols;
      ph[j] = ok ? __ldg(dphi + pm + j) : 0.f;
      rh[j] = ok ? __ldg(rho + pm + j) : 0.f;
    }
    float t[4], eh[4], el[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      t[j] = rh[j] * da[j] * x[j];
      const float e = ph[j] * x[j];
      if (e_lo) { eh[j] = tf32_round(e); el[j] = tf32_round(e - eh[j]); } else { eh[j] = e; el[j] = 0.f; }
    }
    if (full) {
      *reinterpret_cast<float4*>(T + o) = *reinterpret_cast<const float4*>(t);
      *reinterpret_cast<float4*>(e_hi + o) = *reinterpret_cast<const float4*>(eh);
      if (e_lo) *reinterpret_cast<float4*>(e_lo + o) = *reinterpret_cast<const float4*>(el);
    } else {
#pragma unroll
      for (int j = 0; j < 4; ++j)
        if (c4 + j < cols) {
          T[o + j] = t[j];
          e_hi[o + j] = eh[j];
          if (e_lo) e_lo[o + j] = el[j];
        }
    }
  }
}

static inline int64_t pad4l(int64_t x) { return (x + 3) / 4 * 4; }

bool zg_tc_ok(const lip_model* m) {
  static const bool on = !(getenv("LIP_ZGRAD_TC") && atoi(getenv("LIP_ZGRAD_TC")) == 0);
  if (!on || !m->tc_on) return false;
  for (size_t l = 0; l + 1 < m->L.size(); ++l)
    if (m->L[l].act != LIP_OP_TANH && m->L[l].act != LIP_OP_RELU) return false;
  return true;
}

int64_t zg_tc_wld(const lip_model* m) {
  int64_t w = pad4l(m->L[0].in);
  for (auto& L : m->L) w = pad4l(L.out) > w ? pad4l(L.out) : w;
  return w;
}

struct ZTcSizes {
  size_t mlp, keep, eq, small, tmp, total;
};

ZTcSizes zg_tc_sizes(const lip_model* m, int64_t B, int nseg) {
  ZTcSizes z;
  const int nL = (int)m->L.size();
  z.mlp = align_up(mlp_ws_bytes(m, B), 256);
  size_t keep = 0;
  for (int l = 0; l + 1 < nL; ++l) keep += 2 * align_up(sizeof(float) * (size_t)B * m->M * mlp_ld(m, m->L[l].out), 256);
  z.keep = keep;
  z.eq = align_up(sizeof(float) * (size_t)B * m->M * zg_tc_wld(m), 256);         // one of the 10 e / q / X / T buffers
  z.small = align_up(sizeof(float) * (size_t)nseg * B * m->M * m->K, 256);       // dl, Cc, Gf
  z.tmp = align_up(sizeof(float) * (size_t)nseg * B * m->M * m->L[0].in, 256);
  z.total = nseg * (z.mlp + z.keep) + 10 * z.eq + 3 * z.small + z.tmp + 512;
  return z;
}

int zgrad_tc(lip_model* m, int32_t mode, const float* X1, const float* X2, float* out, int64_t B, float scale, int32_t per_probe,
             void* workspace, size_t workspace_bytes, cudaStream_t st) {
  const int nL = (int)m->L.size();
  const int64_t M = m->M;
  const int nseg = mode == LIP_ZGRAD_GGN ? 2 : 1;
  const int64_t nb = nseg * B;
  const ZTcSizes sz = zg_tc_sizes(m, B, nseg);
  if (!workspace || workspace_bytes < sz.total) {
    set_error("lip_zgrad: workspace too small: need %zu bytes, got %zu", sz.total, workspace_bytes);
    return LIP_ERR_WORKSPACE;
  }
  char* base = (char*)align_up((uintptr_t)workspace, 256);
  auto take = [&](size_t bytes) { char* p = base; base += bytes; return (float*)p; };
  float* mlp_ws[2] = {nullptr, nullptr};
  std::vector<float*> keep_hi[2], keep_lo[2];
  for (int sg = 0; sg < nseg; ++sg) {
    mlp_ws[sg] = take(sz.mlp);
    keep_hi[sg].assign(nL, nullptr); keep_lo[sg].assign(nL, nullptr);
    for (int l = 0; l + 1 < nL; ++l) {
      const size_t b = align_up(sizeof(float) * (size_t)B * M * mlp_ld(m, m->L[l].out), 256);
      keep_hi[sg][l] = take(b);
      keep_lo[sg][l] = take(b);
    }
  }
  float* eq[10];
  for (int i = 0; i < 10; ++i) eq[i] = take(sz.eq);
  float* dl = take(sz.small);
  float* Cc = take(sz.small);
  float* Gf = take(sz.small);
  float* tmp = take(sz.tmp);
  const float* Vseg[2] = {X1, mode == LIP_ZGRAD_GGN ? X2 : nullptr};
  const bool classifier = m->model_type == LIP_CLASSIFIER;

  // ---- forward tangent pass = the JVP sweep, every layer's masked tangent kept ----
  const float* vs_hi[2] = {nullptr, nullptr};
  const float* vs_lo[2] = {nullptr, nullptr};
  for (int sg = 0; sg < nseg; ++sg) {
    int rc = mlp_jvp_keep(m, Vseg[sg], B, mlp_ws[sg], sz.mlp, dl + (int64_t)sg * B * M * m->K, keep_hi[sg].data(), keep_lo[sg].data(),
                          &vs_hi[sg], &vs_lo[sg], st);
    if (rc) return rc;
  }

  // ---- output-space rows ----
  float s = scale;
  if (!classifier && (mode == LIP_ZGRAD_WT || mode == LIP_ZGRAD_W)) s *= expf(-0.5f * m->logvar);
  zgrad_rows_kernel<<<(unsigned)ceil_div(B * M, 128), 128, 0, st>>>(mode, classifier ? 1 : 0, dl, X2, m->P, m->S, Cc, Gf, B, M, m->K, s);
  LIP_LAUNCH_CHECK();

  // ---- reverse pass, one segment at a time ----
  float* X = eq[8];
  float* T = eq[9];
  for (int sg = 0; sg < nseg; ++sg) {
    const float* e_hi = Cc + (int64_t)sg * B * M * m->K;
    const float* q_hi = Gf + (int64_t)sg * B * M * m->K;
    const float* e_lo = nullptr;
    const float* q_lo = nullptr;
    int cur_ld = m->K;
    int flip = 0;     // next e / q pairs go to eq[flip .. flip + 3]
    for (int l = nL - 1; l >= 0; --l) {
      const DenseLayer& Ld = m->L[l];
      const bool tc = m->tc_layer[l] != 0;
      if (tc && !e_lo) {   // a tensor-core top layer: re-lay the fp32 rows as padded (hi, lo) pairs
        const int ldp = (int)pad4l(cur_ld);
        int rc = tf32_split(e_hi, cur_ld, eq[flip], eq[flip + 1], ldp, B * M, Ld.out, st);
        if (rc) return rc;
        rc = tf32_split(q_hi, cur_ld, eq[flip + 2], eq[flip + 3], ldp, B * M, Ld.out, st);
        if (rc) return rc;
        e_hi = eq[flip]; e_lo = eq[flip + 1]; q_hi = eq[flip + 2]; q_lo = eq[flip + 3];
        cur_ld = ldp; flip ^= 4;
      }
      const int in_ld = l > 0 ? mlp_ld(m, Ld.in) : Ld.in;
      const bool next_pair = l > 0 && m->tc_layer[l - 1] != 0;
      const int64_t per_in = M * (int64_t)in_ld, per_out = M * (int64_t)cur_ld;
      float* en_hi = eq[flip];
      float* en_lo = next_pair ? eq[flip + 1] : nullptr;
      float* qn_hi = l > 0 ? eq[flip + 2] : tmp + (int64_t)sg * B * M * Ld.in;
      float* qn_lo = next_pair ? eq[flip + 3] : nullptr;
      if (l > 0) {   // X = e_l W_l^T
        if (tc) {
          TcGemmProblem p;
          p.M = M; p.N = Ld.in; p.K = Ld.out; p.batch = B;
          p.A1.hi = e_hi; p.A1.lo = e_lo; p.A1.ld = cur_ld; p.A1.sz = per_out; p.A1.major_k = 1; p.a_batched = 1;
          p.B1.hi = m->W_hi[l]; p.B1.lo = m->W_lo[l]; p.B1.ld = m->W_ld[l]; p.B1.sz = (int64_t)Ld.in * m->W_ld[l]; p.B1.major_k = 1;
          p.b_batched = 0;
          p.C = X; p.c_sz = per_in; p.c_sm = in_ld;
          int rc = gemm_tc(p, st);
          if (rc) return rc;
        } else {
          GemmProblem p;
          p.M = M; p.N = Ld.in; p.K = Ld.out; p.batch = B;
          p.A1 = {e_hi, per_out, cur_ld, 1};
          p.B1 = {m->theta + Ld.woff, 0, 1, Ld.out};
          p.C = X; p.c_sz = per_in; p.c_sm = in_ld;
          int rc = gemm_simt(p, st);
          if (rc) return rc;
        }
        const dim3 rgrid((unsigned)ceil_div(ceil_div(Ld.in, 4), 128), (unsigned)(B * M < 16384 ? B * M : 16384));
        reverse_act_tc_kernel<<<rgrid, 128, 0, st>>>(X, keep_hi[sg][l - 1], tc ? keep_lo[sg][l - 1] : nullptr,   // dA_l has a lo part iff its consumer (layer l) is a tensor-core layer
                                                                    
                                                                     m->dphi[l - 1], m->rho[l - 1], en_hi, en_lo, T, B * M, M, Ld.in, in_ld);
        LIP_LAUNCH_CHECK();
      }
      // q_{l-1} = phi' * (e_l dW_l^T + q_l W_l^T) + T        (l = 0: dZ[b] = e_0 dW_0^T + q_0 W_0^T, no epilogue)
      if (tc) {
        const int64_t ldw = m->W_ld[l];
        TcGemmProblem p;
        p.M = M; p.N = Ld.in; p.K = Ld.out; p.K2 = Ld.out; p.batch = B;
        p.A1.hi = e_hi; p.A1.lo = e_lo; p.A1.ld = cur_ld; p.A1.sz = per_out; p.A1.major_k = 1; p.a_batched = 1;
        p.B1.hi = vs_hi[sg] + B * m->split_off[l]; p.B1.lo = vs_lo[sg] + B * m->split_off[l]; p.B1.ld = ldw;
        p.B1.sz = (int64_t)Ld.in * ldw; p.B1.major_k = 1; p.b_batched = 1;
        p.A2.hi = q_hi; p.A2.lo = q_lo; p.A2.ld = cur_ld; p.A2.sz = per_out; p.A2.major_k = 1; p.a2_batched = 1;
        p.B2.hi = m->W_hi[l]; p.B2.lo = m->W_lo[l]; p.B2.ld = ldw; p.B2.sz = (int64_t)Ld.in * ldw; p.B2.major_k = 1; p.b2_batched = 0;
        p.C = qn_hi; p.C_lo = qn_lo; p.c_sz = per_in; p.c_sm = in_ld;
        if (l > 0) {
          p.epi.mask = m->dphi[l - 1]; p.epi.mask_sm = Ld.in;
          p.epi.add = T; p.epi.add_sz = per_in; p.epi.add_scale = 1.f;
        }
        int rc = gemm_tc(p, st);
        if (rc) return rc;
      } else {
        GemmProblem p;
        p.M = M; p.N = Ld.in; p.K = Ld.out; p.batch = B;
        p.A1 = {e_hi, per_out, cur_ld, 1};
        p.B1 = {Vseg[sg] + Ld.woff, m->D, 1, Ld.out};
        p.A2 = {q_hi, per_out, cur_ld, 1};
        p.B2 = {m->theta + Ld.woff, 0, 1, Ld.out};
        p.K2 = Ld.out;
        p.C = qn_hi; p.c_sz = per_in; p.c_sm = in_ld;
        p.epi.C_lo = qn_lo;
        if (l > 0) {
          p.epi.mask = m->dphi[l - 1]; p.epi.mask_sm = Ld.in;
          p.epi.add = T; p.epi.add_sz = per_in; p.epi.add_scale = 1.f;
        }
        int rc = gemm_simt(p, st);
        if (rc) return rc;
      }
      e_hi = en_hi; e_lo = en_lo; q_hi = qn_hi; q_lo = qn_lo;
      cur_ld = in_ld; flip ^= 4;
    }
  }
  const int64_t per0 = M * (int64_t)m->L[0].in;
  if (!per_probe) {
    batch_sum_kernel<<<blocks(per0), 256, 0, st>>>(tmp, out, per0, nb, 1.f);
  } else if (mode == LIP_ZGRAD_GGN) {
    batch_sum_kernel<<<blocks(B * per0), 256, 0, st>>>(tmp, out, B * per0, 2, 1.f);
  } else {
    batch_sum_kernel<<<blocks(B * per0), 256, 0, st>>>(tmp, out, B * per0, 1, 1.f);
  }
  LIP_LAUNCH_CHECK();
  return LIP_OK;
}

}  // namespace
}  // namespace lip

namespace lip {
int launch_zgrad_rows(int mode, const lip_model* m, const float* dl, const float* X2, float* Cc, float* Gf, int64_t B, float scale,
                      cudaStream_t st) {
  const bool classifier = m->model_type == LIP_CLASSIFIER;
  float s = scale;
  if (!classifier && (mode == LIP_ZGRAD_WT || mode == LIP_ZGRAD_W)) s *= expf(-0.5f * m->logvar);   // as lip_w(t)_apply, factor SQRT
  zgrad_rows_kernel<<<(unsigned)ceil_div(B * m->M, 128), 128, 0, st>>>(mode, classifier ? 1 : 0, dl, X2, m->P, m->S, Cc, Gf, B, m->M,
                                                                      m->K, s);
  LIP_LAUNCH_CHECK();
  return LIP_OK;
}
int launch_batch_sum(const float* x, float* out, int64_t per, int64_t nb, cudaStream_t st) {
  batch_sum_kernel<<<blocks(per), 256, 0, st>>>(x, out, per, nb, 1.f);
  LIP_LAUNCH_CHECK();
  return LIP_OK;
}

// Bind-time: phi''(h_l) and rho_l = phi''/phi' at the bound points (they depend on Z and theta only).  The pre-activations are
// recomputed (the forward pass keeps phi(h) and phi'(h), not h).
int zgrad_prepare(lip_model* m, cudaStream_t st) {
  const int nL = (int)m->L.size();
  m->ddphi.assign(nL > 1 ? nL - 1 : 0, nullptr);
  m->rho.assign(nL > 1 ? nL - 1 : 0, nullptr);
  for (int l = 0; l + 1 < nL; ++l) {
    const DenseLayer& Ld = m->L[l];
    const int64_t n = m->M * (int64_t)Ld.out;
    LIP_CHECK_CUDA(cudaMalloc(&m->ddphi[l], sizeof(float) * (size_t)n + 256));
    LIP_CHECK_CUDA(cudaMalloc(&m->rho[l], sizeof(float) * (size_t)n + 256));
    GemmProblem p;
    p.M = m->M; p.N = Ld.out; p.K = Ld.in; p.batch = 1;
    p.A1 = {m->A[l], 0, Ld.in, 1};
    p.B1 = {m->theta + Ld.woff, 0, Ld.out, 1};
    p.C = m->ddphi[l]; p.c_sz = 0; p.c_sm = Ld.out;
    p.epi.bias = m->theta + Ld.boff; p.epi.bias_sz = 0;
    int rc = gemm_simt(p, st);
    if (rc) return rc;
    act_second_kernel<<<blocks(n), 256, 0, st>>>(m->ddphi[l], n, Ld.act);
    LIP_LAUNCH_CHECK();
    LIP_CHECK_CUDA(cudaMemcpyAsync(m->rho[l], m->ddphi[l], sizeof(float) * (size_t)n, cudaMemcpyDeviceToDevice, st));
    rho_kernel<<<blocks(n), 256, 0, st>>>(m->rho[l], m->dphi[l], n);
    LIP_LAUNCH_CHECK();
  }
  return LIP_OK;
}
}  // namespace lip

using namespace lip;

extern "C" {

size_t lip_zgrad_workspace_bytes(const lip_model* m, int32_t mode, int64_t B) {
  if (!m || !m->bound || B <= 0) return 0;
  if (m->is_resnet) return resnet_zgrad_ws_bytes(m, mode, B);
  if (m->is_cnn) return cnn_zgrad_ws_bytes(m, mode, B);
  if (zg_tc_ok(m)) return zg_tc_sizes(m, B, mode == LIP_ZGRAD_GGN ? 2 : 1).total;
  return zg_bytes(m, mode == LIP_ZGRAD_GGN ? 2 * B : B);
}

int lip_zgrad(lip_model* m, int32_t mode, const float* X1, const float* X2, float* out, int64_t B, float scale,
              int32_t per_probe, void* workspace, size_t workspace_bytes, lip_stream_t stream) {
  LIP_REQUIRE(m && X1 && X2 && out && B > 0, "lip_zgrad: null argument or B <= 0");
  LIP_REQUIRE(mode >= LIP_ZGRAD_GGN && mode <= LIP_ZGRAD_JVP, "lip_zgrad: bad mode %d", mode);
  if (!m->bound) { set_error("lip_zgrad: model not bound"); return LIP_ERR_NOT_BOUND; }
  cudaStream_t st = (cudaStream_t)stream;
  if (m->is_resnet) return resnet_zgrad(m, mode, X1, X2, out, B, scale, per_probe, workspace, workspace_bytes, st);
  if (m->is_cnn) return cnn_zgrad(m, mode, X1, X2, out, B, scale, per_probe, workspace, workspace_bytes, st);
  if (zg_tc_ok(m)) return zgrad_tc(m, mode, X1, X2, out, B, scale, per_probe, workspace, workspace_bytes, st);
  const int nL = (int)m->L.size();
  const int64_t M = m->M;
  const int64_t nb = mode == LIP_ZGRAD_GGN ? 2 * B : B;
  const size_t need = zg_bytes(m, nb);
  if (!workspace || workspace_bytes < need) {
    set_error("lip_zgrad: workspace too small: need %zu bytes, got %zu", need, workspace_bytes);
    return LIP_ERR_WORKSPACE;
  }
  ZWs w;
  zg_carve(m, nb, workspace, &w);
  Seg segs[2];
  int nseg = 1;
  segs[0] = {X1, 0, B};
  if (mode == LIP_ZGRAD_GGN) { segs[1] = {X2, B, B}; nseg = 2; }
  const bool classifier = m->model_type == LIP_CLASSIFIER;

  // ---- forward tangent pass, raw dh_l kept for every hidden layer ----
  float* dA = w.buf[0];   // phi' * dh of the previous layer
  for (int l = 0; l < nL; ++l) {
    const DenseLayer& Ld = m->L[l];
    const bool last = l == nL - 1;
    float* dst = last ? w.dl : w.dh[l];
    for (int sgi = 0; sgi < nseg; ++sgi) {
      const Seg& sg = segs[sgi];
      GemmProblem p;
      p.M = M; p.N = Ld.out; p.K = Ld.in; p.batch = sg.nb;
      p.A1 = {m->A[l], 0, Ld.in, 1};
      p.B1 = {sg.V + Ld.woff, m->D, Ld.out, 1};
      if (l > 0) {
        p.A2 = {dA + sg.b0 * M * Ld.in, M * (int64_t)Ld.in, Ld.in, 1};
        p.B2 = {m->theta + Ld.woff, 0, Ld.out, 1};
        p.K2 = Ld.in;
      }
      p.C = dst + sg.b0 * M * Ld.out; p.c_sz = M * (int64_t)Ld.out; p.c_sm = Ld.out;
      p.epi.bias = sg.V + Ld.boff; p.epi.bias_sz = m->D;
      int rc = gemm_simt(p, st);
      if (rc) return rc;
    }
    if (!last) {
      const int64_t per = M * Ld.out;
      mask_mul_kernel<<<blocks(nb * per), 256, 0, st>>>(w.dh[l], m->dphi[l], dA, per, nb * per);
      LIP_LAUNCH_CHECK();
    }
  }

  // ---- output-space rows: e_L = c, q_L = g ----
  float s = scale;
  if (!classifier && (mode == LIP_ZGRAD_WT || mode == LIP_ZGRAD_W)) s *= expf(-0.5f * m->logvar);   // as lip_w(t)_apply, factor SQRT
  zgrad_rows_kernel<<<(unsigned)ceil_div(B * M, 128), 128, 0, st>>>(mode, classifier ? 1 : 0, w.dl, X2, m->P, m->S, w.Cc, w.Gf, B, M,
                                                                   m->K, s);
  LIP_LAUNCH_CHECK();

  // ---- reverse pass ----
  const float* e = w.Cc;
  const float* q = w.Gf;
  int flip = 0;   // e/q of the next layer go to buf[flip], buf[flip + 1]; then flip ^= 2
  float* X = w.buf[4];
  float* T = w.buf[5];
  for (int l = nL - 1; l >= 0; --l) {
    const DenseLayer& Ld = m->L[l];
    const int64_t per_in = M * Ld.in, per_out = M * Ld.out;
    if (l > 0) {
      {  // X = e_l W_l^T
        GemmProblem p;
        p.M = M; p.N = Ld.in; p.K = Ld.out; p.batch = nb;
        p.A1 = {e, per_out, Ld.out, 1};
        p.B1 = {m->theta + Ld.woff, 0, 1, Ld.out};
        p.C = X; p.c_sz = per_in; p.c_sm = Ld.in;
        int rc = gemm_simt(p, st);
        if (rc) return rc;
      }
      float* e_next = w.buf[flip];
      float* q_next = w.buf[flip + 1];
      reverse_act_kernel<<<blocks(nb * per_in), 256, 0, st>>>(X, w.dh[l - 1], m->dphi[l - 1], m->ddphi[l - 1], e_next, T, per_in,
                                                              nb * per_in);
      LIP_LAUNCH_CHECK();
      for (int sgi = 0; sgi < nseg; ++sgi) {   // q_{l-1} = phi' * (e_l dW_l^T + q_l W_l^T) + T
        const Seg& sg = segs[sgi];
        GemmProblem p;
        p.M = M; p.N = Ld.in; p.K = Ld.out; p.batch = sg.nb;
        p.A1 = {e + sg.b0 * per_out, per_out, Ld.out, 1};
        p.B1 = {sg.V + Ld.woff, m->D, 1, Ld.out};
        p.A2 = {q + sg.b0 * per_out, per_out, Ld.out, 1};
        p.B2 = {m->theta + Ld.woff, 0, 1, Ld.out};
        p.K2 = Ld.out;
        p.C = q_next + sg.b0 * per_in; p.c_sz = per_in; p.c_sm = Ld.in;
        p.epi.mask = m->dphi[l - 1]; p.epi.mask_sm = Ld.in;
        p.epi.add = T + sg.b0 * per_in; p.epi.add_sz = per_in; p.epi.add_scale = 1.f;
        int rc = gemm_simt(p, st);
        if (rc) return rc;
      }
      e = e_next; q = q_next; flip ^= 2;
    } else {
      float* dst = per_probe && mode != LIP_ZGRAD_GGN ? out : X;
      for (int sgi = 0; sgi < nseg; ++sgi) {   // dZ[b] = e_0 dW_0^T + q_0 W_0^T
        const Seg& sg = segs[sgi];
        GemmProblem p;
        p.M = M; p.N = Ld.in; p.K = Ld.out; p.batch = sg.nb;
        p.A1 = {e + sg.b0 * per_out, per_out, Ld.out, 1};
        p.B1 = {sg.V + Ld.woff, m->D, 1, Ld.out};
        p.A2 = {q + sg.b0 * per_out, per_out, Ld.out, 1};
        p.B2 = {m->theta + Ld.woff, 0, 1, Ld.out};
        p.K2 = Ld.out;
        p.C = dst + sg.b0 * per_in; p.c_sz = per_in; p.c_sm = Ld.in;
        int rc = gemm_simt(p, st);
        if (rc) return rc;
      }
      if (!per_probe) {
        batch_sum_kernel<<<blocks(per_in), 256, 0, st>>>(X, out, per_in, nb, 1.f);
        LIP_LAUNCH_CHECK();
      } else if (mode == LIP_ZGRAD_GGN) {
        // per probe: the ubar-side and v-side halves of probe b add up: out[b] = X[b] + X[B + b]
        batch_sum_kernel<<<blocks(B * per_in), 256, 0, st>>>(X, out, B * per_in, 2, 1.f);
        LIP_LAUNCH_CHECK();
      }
    }
  }
  return LIP_OK;
}

}  // extern "C"
