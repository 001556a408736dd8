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
#include <new>

#include "lip_common.cuh"

using namespace lip;

struct DenseLayer {
  int in = 0, out = 0;
  int64_t boff = 0, woff = 0;
  int act = -1;  // activation applied to this layer's output (-1: none / last layer)
};

struct lip_model {
  std::vector<DenseLayer> L;
  int model_type = LIP_CLASSIFIER;
  int64_t D = 0;
  int K = 0;
  int maxw = 0;  // widest layer output
  // bound state
  bool bound = false;
  int64_t M = 0;
  const float* theta = nullptr;
  float logvar = 0.f;
  std::vector<float*> A;     // A[l]: input of layer l, [M, in_l]   (A[0] = Z)
  std::vector<float*> dphi;  // dphi[l]: phi'(h_l) at the output of layer l (l < nL-1), [M, out_l]
  float* logits = nullptr;   // [M, K]
  float* P = nullptr;        // softmax(logits)
  float* S = nullptr;        // sqrt(P)
  int use_tc = -1;           // -1 auto
  // tcgen05 operands: TF32 hi/lo splits in padded buffers (ld multiple of 32)
  std::vector<float*> A_hi, A_lo, W_hi, W_lo;
  std::vector<int64_t> A_ld, W_ld;

  void free_cache() {
    for (auto p : A) if (p) cudaFree(p);
    for (auto p : dphi) if (p) cudaFree(p);
    for (auto p : A_hi) if (p) cudaFree(p);
    for (auto p : A_lo) if (p) cudaFree(p);
    for (auto p : W_hi) if (p) cudaFree(p);
    for (auto p : W_lo) if (p) cudaFree(p);
    A.clear(); dphi.clear(); A_hi.clear(); A_lo.clear(); W_hi.clear(); W_lo.clear(); A_ld.clear(); W_ld.clear();
    if (logits) cudaFree(logits);
    if (P) cudaFree(P);
    if (S) cudaFree(S);
    logits = P = S = nullptr;
    bound = false;
  }
};

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

// gb[b][j] = scale * sum_m Delta[b][m][j] + add_scale * add[b][j]
__global__ void bias_grad_kernel(const float* __restrict__ Delta, int64_t M, int n, float* __restrict__ out,
                                 int64_t out_sz, float scale, const float* __restrict__ add, int64_t add_sz,
                                 float add_scale) {
  const int b = blockIdx.y;
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= n) return;
  const float* d = Delta + (int64_t)b * M * n + j;
  float acc = 0.f;
  int64_t m = 0;
  for (; m + 4 <= M; m += 4) {
    float a0 = d[(m + 0) * n], a1 = d[(m + 1) * n], a2 = d[(m + 2) * n], a3 = d[(m + 3) * n];
    acc += (a0 + a1) + (a2 + a3);
  }
  for (; m < M; ++m) acc += d[m * n];
  float v = scale * acc;
  if (add) v += add_scale * add[(int64_t)b * add_sz + j];
  out[(int64_t)b * out_sz + j] = v;
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

struct Workspace {
  float* buf[2];
  size_t per_buf;  // floats
};

size_t ws_bytes(const lip_model* m, int64_t B) {
  size_t per = (size_t)B * (size_t)m->M * (size_t)m->maxw;
  per = align_up(per, 64);
  return 2 * per * sizeof(float) + 256;
}

int carve(const lip_model* m, int64_t B, void* ws, size_t bytes, Workspace* w) {
  size_t need = ws_bytes(m, B);
  if (bytes < need || ws == nullptr) {
    set_error("workspace too small: need %zu bytes, got %zu", need, bytes);
    return LIP_ERR_WORKSPACE;
  }
  uintptr_t base = align_up((uintptr_t)ws, 256);
  size_t per = align_up((size_t)B * (size_t)m->M * (size_t)m->maxw, 64);
  w->buf[0] = (float*)base;
  w->buf[1] = w->buf[0] + per;
  w->per_buf = per;
  return LIP_OK;
}

// ---- JVP sweep: V[B,D] -> dlogits written to `dst` ([B,M,K], contiguous).  Intermediates ping-pong in ws.
int jvp_sweep(lip_model* m, const float* V, int64_t B, const Workspace& w, float* dst, float out_scale,
              cudaStream_t st) {
  const int nL = (int)m->L.size();
  const float* prev = nullptr;
  for (int l = 0; l < nL; ++l) {
    const DenseLayer& Ld = m->L[l];
    const bool last = (l == nL - 1);
    float* out = last ? dst : w.buf[l & 1];
    GemmProblem p;
    p.M = m->M; p.N = Ld.out; p.K = Ld.in; p.batch = B;
    p.A1 = {m->A[l], 0, Ld.in, 1};
    p.B1 = {V + Ld.woff, m->D, Ld.out, 1};
    if (l > 0) {
      p.A2 = {prev, m->M * (int64_t)Ld.in, Ld.in, 1};
      p.B2 = {m->theta + Ld.woff, 0, Ld.out, 1};
      p.K2 = Ld.in;
    }
    p.C = out; p.c_sz = m->M * (int64_t)Ld.out; p.c_sm = Ld.out;
    p.epi.bias = V + Ld.boff; p.epi.bias_sz = m->D;
    if (!last) { p.epi.mask = m->dphi[l]; p.epi.mask_sm = Ld.out; }
    // NB epilogue order is scale*acc + bias, then mask: out_scale only applies on the last layer (no mask),
    // where scale*(acc) + bias would mis-scale the bias -> apply out_scale separately when != 1.
    int rc = gemm_simt(p, st);
    if (rc) return rc;
    prev = out;
  }
  (void)out_scale;
  return LIP_OK;
}

// ---- VJP sweep: Delta_L in `delta` ([B,M,K]); writes out[B,D] = scale * J^T delta + add_scale * add.
// `delta` lives in (or is copied to) a workspace buffer; the other buffer is used for ping-pong.
int vjp_sweep(lip_model* m, float* delta, float* other, int64_t B, float* out, float scale, const float* add,
              float add_scale, cudaStream_t st) {
  const int nL = (int)m->L.size();
  float* cur = delta;
  float* nxt = other;
  for (int l = nL - 1; l >= 0; --l) {
    const DenseLayer& Ld = m->L[l];
    {  // weight gradient: [in x out] = A_l^T [in x M] * Delta [M x out]
      GemmProblem p;
      p.M = Ld.in; p.N = Ld.out; p.K = m->M; p.batch = B;
      p.A1 = {m->A[l], 0, 1, Ld.in};
      p.B1 = {cur, m->M * (int64_t)Ld.out, Ld.out, 1};
      p.C = out + Ld.woff; p.c_sz = m->D; p.c_sm = Ld.out;
      p.epi.scale = scale;
      if (add) { p.epi.add = add + Ld.woff; p.epi.add_sz = m->D; p.epi.add_scale = add_scale; }
      int rc = gemm_simt(p, st);
      if (rc) return rc;
    }
    {  // bias gradient
      dim3 grid((unsigned)ceil_div(Ld.out, 128), (unsigned)B);
      bias_grad_kernel<<<grid, 128, 0, st>>>(cur, m->M, Ld.out, out + Ld.boff, m->D, scale,
                                             add ? add + Ld.boff : nullptr, m->D, add_scale);
      LIP_LAUNCH_CHECK();
    }
    if (l > 0) {  // Delta_{l-1} = (Delta_l W_l^T) * phi'_{l-1}
      GemmProblem p;
      p.M = m->M; p.N = Ld.in; p.K = Ld.out; p.batch = B;
      p.A1 = {cur, m->M * (int64_t)Ld.out, Ld.out, 1};
      p.B1 = {m->theta + Ld.woff, 0, 1, Ld.out};
      p.C = nxt; p.c_sz = m->M * (int64_t)Ld.in; p.c_sm = Ld.in;
      p.epi.mask = m->dphi[l - 1]; p.epi.mask_sm = Ld.in;
      int rc = gemm_simt(p, st);
      if (rc) return rc;
      float* t = cur; cur = nxt; nxt = t;
    }
  }
  return LIP_OK;
}

}  // namespace

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
  m->free_cache();
  delete m;
  return LIP_OK;
}

int64_t lip_model_num_params(const lip_model* m) { return m ? m->D : -1; }
int64_t lip_model_num_outputs(const lip_model* m) { return m ? m->K : -1; }
int64_t lip_model_num_points(const lip_model* m) { return (m && m->bound) ? m->M : -1; }

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
  m->bound = true;
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
  return ws_bytes(m, B);
}

int lip_ggn_vp(lip_model* m, const float* V, float* out, int64_t B, float recal, float alpha, void* workspace,
               size_t workspace_bytes, lip_stream_t stream) {
  LIP_REQUIRE(m && V && out && B > 0, "lip_ggn_vp: null argument or B <= 0");
  LIP_REQUIRE(V != out, "lip_ggn_vp: in-place operation is not supported");
  if (!m->bound) { set_error("lip_ggn_vp: model not bound"); return LIP_ERR_NOT_BOUND; }
  cudaStream_t st = (cudaStream_t)stream;
  Workspace w;
  int rc = carve(m, B, workspace, workspace_bytes, &w);
  if (rc) return rc;
  const int nL = (int)m->L.size();
  // dlogits land in the buffer the last hidden layer did NOT use, so the VJP can ping-pong from it
  float* dl = w.buf[(nL - 1) & 1];
  rc = jvp_sweep(m, V, B, w, dl, 1.f, st);
  if (rc) return rc;
  if (m->model_type == LIP_CLASSIFIER) {
    rc = launch_factor(dl, dl, m, B, 0, 1.f, st);
    if (rc) return rc;
  }
  return vjp_sweep(m, dl, w.buf[nL & 1], B, out, recal, alpha != 0.f ? V : nullptr, alpha, st);
}

int lip_wt_apply(lip_model* m, const float* V, float* out, int64_t B, float scale, int32_t factor,
                 void* workspace, size_t workspace_bytes, lip_stream_t stream) {
  LIP_REQUIRE(m && V && out && B > 0, "lip_wt_apply: null argument or B <= 0");
  if (!m->bound) { set_error("lip_wt_apply: model not bound"); return LIP_ERR_NOT_BOUND; }
  cudaStream_t st = (cudaStream_t)stream;
  Workspace w;
  int rc = carve(m, B, workspace, workspace_bytes, &w);
  if (rc) return rc;
  rc = jvp_sweep(m, V, B, w, out, 1.f, st);
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
  Workspace w;
  int rc = carve(m, B, workspace, workspace_bytes, &w);
  if (rc) return rc;
  float s = scale;
  if (factor == LIP_FACTOR_SQRT && m->model_type == LIP_CLASSIFIER) {
    rc = launch_factor(U, w.buf[0], m, B, 2, 1.f, st);
  } else {
    if (factor == LIP_FACTOR_SQRT) s *= expf(-0.5f * m->logvar);
    int64_t n = B * m->M * m->K;
    scale_copy_kernel<<<(unsigned)ceil_div(n, 256), 256, 0, st>>>(U, w.buf[0], n, 1.f);
    LIP_LAUNCH_CHECK();
    rc = LIP_OK;
  }
  if (rc) return rc;
  return vjp_sweep(m, w.buf[0], w.buf[1], B, out, s, add, add_scale, st);
}

size_t lip_gram_workspace_bytes(const lip_model* m, int64_t block) {
  if (!m || !m->bound || block <= 0) return 0;
  size_t d = (size_t)m->M * m->K;
  return ws_bytes(m, block) + align_up(sizeof(float) * (size_t)block * d, 256) +
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

}  // extern "C"
