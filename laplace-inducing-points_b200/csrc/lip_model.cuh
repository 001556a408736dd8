// lip_model.cuh — the model handle shared by the MLP path (lip_model.cu) and the conv stage-program path (lip_cnn.cu).
#pragma once
#include <vector>

#include "lip_common.cuh"

struct DenseLayer {
  int in = 0, out = 0;
  int64_t boff = 0, woff = 0;
  int act = -1;  // activation applied to this layer's output (-1: none / last layer)
};

// One affine stage of a conv program: CONV2D (as a GEMM over im2col patches) or DENSE, + activation + optional 2x2 avg-pool.
struct ConvStage {
  int type = 0;          // 0 dense, 1 conv
  int act = -1;          // activation after the affine op (-1 none)
  int pool = 0;          // 1: nn.avg_pool 2x2 after the activation
  int Hi = 1, Wi = 1, cin = 0, pad = 0, kh = 1, kw = 1, Ho = 1, Wo = 1, cout = 0, Hp = 1, Wp = 1;
  int64_t boff = 0, woff = 0;
  int P = 1;             // GEMM rows per point: Ho*Wo (1 for dense)
  int Kc = 0;            // contraction length: kh*kw*cin (in_features for dense)
  float* Aop = nullptr;  // bound cache: [M*P, Kc] im2col patches (conv) or the stage input (dense)
  float* dphi = nullptr; // bound cache: [M*P, cout] activation derivative (null for the last stage)
  float* Xin = nullptr;  // bound cache (conv): [M, Hi, Wi, cin] stage input image, read by the fused stage kernels (lip_cnn_fused.cu)
  int64_t out_per_point() const { return pool ? (int64_t)Hp * Wp * cout : (int64_t)P * cout; }
};

// One conv + BatchNorm(eval) [+ residual add] [+ relu] unit of a residual network (ResNet1M, src/scalemodels.py:70-157):
//   y = relu?( scale * xhat + beta + skip ),  xhat = (conv(x, W) - mean) * rsqrt(var + eps)
struct ConvBN {
  int Hi = 0, Wi = 0, cin = 0, kh = 0, kw = 0, stride = 1, pad_h = 0, pad_w = 0, Ho = 0, Wo = 0, cout = 0;
  int64_t woff = 0, scale_off = 0, beta_off = 0;   // offsets in theta: kernel[kh,kw,cin,cout], BN scale[c], BN bias[c]
  int64_t stats_off = 0;                           // offset of [mean(c), var(c)] in the bn-stats array
  int relu = 0;
  int src = -2, dst = 0, skip = -1;                // tangent-buffer slots (0..3); src -2 = the network input (data)
  int accumulate = 0;                              // VJP: the cotangent of `src` already holds another branch's contribution
  float* Xin = nullptr;    // [M, Hi, Wi, cin] cached input activation (the conv GEMMs gather their patches from it)
  float* Wt = nullptr;     // [kh*kw*cout, cin] kernel re-laid for the delta back-propagation (transposed conv) GEMM
  float* xhat = nullptr;   // [M*Ho*Wo, cout]
  float* mask = nullptr;   // [M*Ho*Wo, cout] relu' of the unit's pre-activation (null without relu)
  float* g = nullptr;      // [cout] scale * rsqrt(var + eps)
  // tcgen05 implicit-GEMM path (lip_conv_tc.cu): TF32 (hi, lo) splits of the cached image and of the two kernel layouts
  bool tc = false;
  float *Xh = nullptr, *Xl = nullptr, *Wh = nullptr, *Wl = nullptr, *Wth = nullptr, *Wtl = nullptr;
  int64_t P() const { return (int64_t)Ho * Wo; }
  int64_t Kc() const { return (int64_t)kh * kw * cin; }
  // gather descriptors (lip_common.cuh: ConvGather) of this conv for the three GEMM roles
  lip::ConvGather gather(int mode) const {
    lip::ConvGather cg;
    cg.mode = mode; cg.Hi = Hi; cg.Wi = Wi; cg.C = (mode == 3) ? cout : cin; cg.pad_h = pad_h; cg.pad_w = pad_w;
    cg.stride = stride; cg.kh = kh; cg.kw = kw; cg.Ho = Ho; cg.Wo = Wo;
    return cg;
  }
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
  std::vector<float*> ddphi; // ddphi[l]: phi''(h_l), and rho[l] = phi''/phi' (0 where phi' = 0): second-order factors of lip_zgrad
  std::vector<float*> rho;
  float* logits = nullptr;   // [M, K]
  float* P = nullptr;        // softmax(logits)
  float* S = nullptr;        // sqrt(P)
  int use_tc = -1;           // -1 auto (tcgen05 when the device is sm_100), 0 SIMT only, 1 tcgen05
  bool tc_on = false;        // decided at bind time
  // tcgen05 operands: TF32 hi/lo splits of the cached activations and of the weights (ld padded to 4)
  std::vector<float*> A_hi, A_lo, W_hi, W_lo;
  std::vector<int64_t> A_ld, W_ld;
  std::vector<char> tc_layer;   // layer l runs its three GEMMs on the tensor cores
  int64_t max_split = 0;        // max over tc layers of in * ldw (floats per probe of the split tangent block)
  int64_t sum_split = 0;        // sum over tc layers of in * ldw
  std::vector<int64_t> split_off;   // per layer: float offset (per probe) of its block inside the split buffers
  int* lo_nz = nullptr;             // [layers] device flags: the probe block of layer l has a non-zero TF32 lo part
  // side stream that runs the probe-block TF32 splits concurrently with the (compute-bound) JVP GEMMs
  cudaStream_t side = nullptr;
  cudaEvent_t ev_fork = nullptr;
  std::vector<cudaEvent_t> ev_split;   // [layer * SPLIT_CHUNKS + chunk]
  static constexpr int SPLIT_CHUNKS = 4;
  int ldmax = 0;                // widest padded intermediate row

  // ---- conv stage programs (LeNet5, src/scalemodels.py:11-49): see lip_cnn.cu ----
  bool is_cnn = false;
  int in_h = 0, in_w = 0, in_c = 0;      // NHWC input image
  std::vector<ConvStage> CS;   // stages of a conv program
  float* cnn_tmp_out = nullptr;          // bind-time scratch
  float* cnn_tmp_x = nullptr;
  // the trailing DENSE stages of a conv stage program as a dense program of their own (absolute offsets into the same flat parameter
  // vector), bound at the features the conv stages produce: its sweeps put LeNet5's 400 -> 120 -> 84 layers on the tcgen05 path
  lip_model* tail = nullptr;
  int tail_first = -1;                   // index in CS of the first stage the tail covers
  bool tail_on = false;                  // decided at bind time (at least one tail layer runs on the tensor cores)

  // ---- residual conv programs (ResNet1M): see lip_resnet.cu ----
  bool is_resnet = false;
  std::vector<ConvBN> RB;                // stem + blocks, in forward order
  int rn_last_slot = 0, rn_H = 0, rn_W = 0, rn_C = 0;   // the tensor fed to the global mean
  int64_t rn_dense_boff = 0, rn_dense_woff = 0;
  float* rn_mean_act = nullptr;          // [M, rn_C] cached global-mean activations (the head's A operand)
  float* rn_stats = nullptr;             // device copy of the BatchNorm running statistics, [sum 2*c]
  int64_t rn_nstats = 0;
  int64_t rn_slot_elems = 0;             // per-point elements of the largest activation tensor
  bool rn_pairs = false;                 // every unit but the stem is on the tcgen05 path: JVP tangents live as TF32 (hi, lo) pairs
  void free_resnet_cache();
  void free_cnn_cache();
  void free_cache() {
    free_cnn_cache();
    free_resnet_cache();
    for (auto p : A) if (p) cudaFree(p);
    for (auto p : dphi) if (p) cudaFree(p);
    for (auto p : ddphi) if (p) cudaFree(p);
    for (auto p : rho) if (p) cudaFree(p);
    ddphi.clear(); rho.clear();
    for (auto p : A_hi) if (p) cudaFree(p);
    for (auto p : A_lo) if (p) cudaFree(p);
    for (auto p : W_hi) if (p) cudaFree(p);
    for (auto p : W_lo) if (p) cudaFree(p);
    A.clear(); dphi.clear(); A_hi.clear(); A_lo.clear(); W_hi.clear(); W_lo.clear(); A_ld.clear(); W_ld.clear();
    tc_layer.clear(); tc_on = false; max_split = 0; sum_split = 0; split_off.clear();
    if (logits) cudaFree(logits);
    if (P) cudaFree(P);
    if (S) cudaFree(S);
    logits = P = S = nullptr;
    bound = false;
  }
};


namespace lip {
// small shared launchers (defined in lip_model.cu)
int launch_factor(const float* in, float* out, const lip_model* m, int64_t B, int mode, float scale, cudaStream_t st);
// gb[b][j] = scale * sum_{r<rows} Delta[b][r][j] (+ Delta_lo) + add_scale * add[b][j];  Delta rows have stride ld
int launch_bias_grad(const float* Delta, const float* Delta_lo, int64_t rows, int n, int64_t ld, int64_t B, float* out,
                     int64_t out_sz, float scale, const float* add, int64_t add_sz, float add_scale, cudaStream_t st);
int launch_scale_copy(const float* in, float* out, int64_t n, float scale, cudaStream_t st);
int launch_softmax(const float* logits, float* P, float* S, int64_t M, int K, cudaStream_t st);

// pieces of lip_zgrad shared with the conv stage programs (lip_zgrad.cu)
int launch_zgrad_rows(int mode, const lip_model* m, const float* dl, const float* X2, float* Cc, float* Gf, int64_t B, float scale,
                      cudaStream_t st);
int launch_batch_sum(const float* x, float* out, int64_t per, int64_t nb, cudaStream_t st);
// gradients with respect to Z for conv stage programs with relu activations (lip_cnn.cu)
size_t cnn_zgrad_ws_bytes(const lip_model* m, int32_t mode, int64_t B);
int cnn_zgrad(lip_model* m, int32_t mode, const float* X1, const float* X2, float* out, int64_t B, float scale, int32_t per_probe,
              void* ws, size_t bytes, cudaStream_t st);
// bind-time part of lip_zgrad (lip_zgrad.cu): phi'' and phi''/phi' at the bound points of a dense program
int zgrad_prepare(lip_model* m, cudaStream_t st);
// MLP sweep pieces shared with lip_zgrad.cu (defined in lip_model.cu)
size_t mlp_ws_bytes(const lip_model* m, int64_t B);
// a dense program used as the tail of a conv stage program (lip_cnn.cu): built from the stages' offsets, swept with a given input
// tangent T0 [B, M, in0] (JVP) and returning the input cotangent [B, M, in0] (VJP).  Workspace: mlp_tail_ws_bytes.
lip_model* mlp_make_tail(const std::vector<ConvStage>& stages, int first, int model_type, int64_t D);
size_t mlp_tail_ws_bytes(const lip_model* tail, int64_t B);
int mlp_tail_jvp(lip_model* tail, const float* V, int64_t ldv, const float* T0, int64_t B, void* ws, size_t bytes, float* dlogits,
                 cudaStream_t st);
int mlp_tail_vjp(lip_model* tail, const float* dl, int64_t B, void* ws, size_t bytes, float* out, int64_t ldo, float scale,
                 const float* add, int64_t lda, float add_scale, float* cot_in, cudaStream_t st);
int mlp_ld(const lip_model* m, int width);
int mlp_jvp_keep(lip_model* m, const float* V, int64_t B, void* ws, size_t bytes, float* dl, float* const* keep_hi,
                 float* const* keep_lo, const float** vs_hi, const float** vs_lo, cudaStream_t st);

// NHWC im2col / col2im (lip_cnn.cu): patches [MZ*Ho*Wo, kh*kw*C], column order (dy, dx, c) = flax HWIO kernel rows
int im2col(const float* in, float* out, int64_t MZ, int Hi, int Wi, int C, int pad_h, int pad_w, int stride, int kh, int kw,
           int Ho, int Wo, cudaStream_t st);
int col2im(const float* col, float* tin, int64_t MZ, int Hi, int Wi, int C, int pad_h, int pad_w, int stride, int kh, int kw,
           int Ho, int Wo, int accumulate, cudaStream_t st);

// conv stage-program path (lip_cnn.cu)
int cnn_parse(lip_model* m, const lip_layer_desc* layers, int32_t n_layers, int64_t num_params);
int cnn_bind(lip_model* m, const float* theta, const float* Z, int64_t M, cudaStream_t st);
size_t cnn_ws_bytes(const lip_model* m, int64_t B);
int cnn_ggn_vp(lip_model* m, const float* V, float* out, int64_t B, float recal, float alpha, void* ws, size_t bytes,
               cudaStream_t st);
int cnn_wt_apply(lip_model* m, const float* V, float* out, int64_t B, float scale, int32_t factor, void* ws, size_t bytes,
                 cudaStream_t st);
int cnn_w_apply(lip_model* m, const float* U, float* out, int64_t B, float scale, int32_t factor, const float* add,
                float add_scale, void* ws, size_t bytes, cudaStream_t st);
// fused conv + mask + pool stage kernels for LeNet5-shaped stages (lip_cnn_fused.cu); LIP_CNN_FUSE=0 keeps the im2col + GEMM path
bool cnn_stage_fusable(const lip_model* m, int stage);
// out[B, M, Hp, Wp, cout] = avgpool2(phi' * (conv(Xin, dW[b]) + conv(T[b], W) + db[b]));  T = nullptr at stage 0
int cnn_fused_jvp(const lip_model* m, int stage, const float* V, int64_t ldv, const float* T, float* out, int64_t B, cudaStream_t st);
// tin: gradient w.r.t. the pooled stage output; writes the stage's kernel / bias gradient blocks of out[B, D] (scale, + add_scale * add)
// and, for stage > 0, gin = the gradient w.r.t. the stage input [B, M, Hi, Wi, cin]
int cnn_fused_vjp(const lip_model* m, int stage, const float* tin, float* gin, float* out, int64_t B, float scale, const float* add,
                  float add_scale, float* scratch, int64_t scratch_elems, cudaStream_t st);
// ResNet1M stem (3 -> 32 channels, 3x3, 32x32): conv JVP + BatchNorm-JVP in one kernel (lip_cnn_fused.cu)
bool resnet_stem_fusable(const ConvBN& u);
int resnet_stem_jvp(const lip_model* m, const ConvBN& u, const float* V, int64_t ldv, float* out, float* out_lo, int64_t B,
                    cudaStream_t st);
// residual conv programs (lip_resnet.cu)
int resnet_parse(lip_model* m, const lip_layer_desc* layers, int32_t n_layers, int64_t num_params);
int resnet_bind(lip_model* m, const float* theta, const float* Z, int64_t M, cudaStream_t st);
size_t resnet_ws_bytes(const lip_model* m, int64_t B);
int resnet_ggn_vp(lip_model* m, const float* V, float* out, int64_t B, float recal, float alpha, void* ws, size_t bytes,
                  cudaStream_t st);
int resnet_wt_apply(lip_model* m, const float* V, float* out, int64_t B, float scale, int32_t factor, void* ws, size_t bytes,
                    cudaStream_t st);
int resnet_w_apply(lip_model* m, const float* U, float* out, int64_t B, float scale, int32_t factor, const float* add,
                   float add_scale, void* ws, size_t bytes, cudaStream_t st);
// gradients with respect to Z for residual programs (lip_resnet.cu)
size_t resnet_zgrad_ws_bytes(const lip_model* m, int32_t mode, int64_t B);
int resnet_zgrad(lip_model* m, int32_t mode, const float* X1, const float* X2, float* out, int64_t B, float scale, int32_t per_probe,
                 void* ws, size_t bytes, cudaStream_t st);
}  // namespace lip
