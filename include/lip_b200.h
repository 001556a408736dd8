/*
 * lip_b200.h — C ABI of liblip_b200.so: the B200-native (sm_100a) matrix-free linearized-Laplace hot path.
 *
 * Every entry point takes plain pointers and sizes (no torch / JAX types), enqueues its work on the
 * caller's CUDA stream, performs no host synchronisation and no allocation on the hot calls (scratch is
 * caller-owned), and returns 0 (LIP_OK) or a negative lip_status; lip_last_error() gives the message.
 * Device pointers are fp32 unless stated.  Flat parameter vectors use the reference's layout
 * (/root/reference/src/utils.py:12-17 flatten_nn_params == ravel_pytree: sorted keys, bias before kernel,
 * Dense kernel [in,out] row-major).
 *
 * Each function cites the reference interface it replaces (paths relative to /root/reference).
 * The reference is pure Python/JAX: these are the symbols a jax.ffi / XLA custom-call shim (or the ctypes
 * loader shipped in laplace-inducing-points_b200/_cabi.py) binds; see INTEGRATION.md.
 */
#ifndef LIP_B200_H_
#define LIP_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct lip_model lip_model;   /* opaque: layer program + activation cache at the bound points */
typedef void* lip_stream_t;           /* a cudaStream_t */

typedef enum {
  LIP_OK = 0,
  LIP_ERR_INVALID = -1,      /* bad argument / shape / unsupported model (-> ValueError / ffi::InvalidArgument) */
  LIP_ERR_CUDA = -2,         /* CUDA runtime failure (-> RuntimeError / ffi::Internal) */
  LIP_ERR_WORKSPACE = -3,    /* caller workspace too small */
  LIP_ERR_UNSUPPORTED = -4,  /* no sm_100a device / feature not built */
  LIP_ERR_NOT_BOUND = -5
} lip_status;

/* layer program ops.  Dense/activation chains are models M1/M2 of SURVEY.md 8a (src/toymodels.py:4-37,
 * src/scalemodels.py:52-67); INPUT/ZEROPAD/CONV2D/AVGPOOL2/FLATTEN add model M3, LeNet5 (src/scalemodels.py:11-49).
 * Images are NHWC, conv kernels HWIO (flax), stride 1, VALID after the explicit ZEROPAD. */
typedef enum {
  LIP_OP_DENSE = 0,      /* y = x @ kernel[in,out] + bias[out]   (flax nn.Dense) */
  LIP_OP_TANH = 1,
  LIP_OP_GELU_TANH = 2,  /* flax nn.gelu(approximate=True) */
  LIP_OP_RELU = 3,
  LIP_OP_CONV2D = 4,     /* y = conv(x, kernel[kh,kw,cin,cout]) + bias[cout], stride 1, VALID (flax nn.Conv) */
  LIP_OP_AVGPOOL2 = 5,   /* nn.avg_pool(window (2,2), strides (2,2)) */
  LIP_OP_ZEROPAD = 6,    /* jnp.pad(x, ((0,0),(p,p),(p,p),(0,0))) */
  LIP_OP_FLATTEN = 7,    /* x.reshape(batch, -1) of an NHWC image (a no-op on the memory layout) */
  LIP_OP_INPUT = 8,      /* declares an image input: kh = height, kw = width, in_features = channels (first op) */
  /* residual networks (model M4, ResNet1M: src/scalemodels.py:70-157).  A program is
   *   INPUT, CONV2D, BATCHNORM, RELU,
   *   { RES_SAVE, CONV2D, BATCHNORM, RELU, CONV2D, BATCHNORM, [RES_CONV2D, RES_BATCHNORM,] RES_ADD, RELU }*,
   *   GLOBAL_MEAN, DENSE
   * with bias-free SAME-padded convs (pad = rows/columns of zeros BEFORE the image, flax/XLA SAME; stride 1 or 2). */
  LIP_OP_BATCHNORM = 9,      /* eval mode (ggn.py:52): in_features = channels, bias_offset = bias[c], kernel_offset = scale[c]
                                in theta; running mean/var come from lip_model_set_bn_stats, in program order */
  LIP_OP_RES_SAVE = 10,      /* remember the current tensor as the block's residual */
  LIP_OP_RES_CONV2D = 11,    /* CONV2D applied to the saved residual (the 1x1 strided shortcut) */
  LIP_OP_RES_BATCHNORM = 12, /* BATCHNORM applied to the saved residual */
  LIP_OP_RES_ADD = 13,       /* x = x + residual */
  LIP_OP_GLOBAL_MEAN = 14    /* jnp.mean(x, axis=(1, 2)) */
} lip_op;

typedef struct {
  int32_t op;            /* lip_op */
  int32_t in_features;   /* DENSE: in; CONV2D: cin; INPUT: channels */
  int32_t out_features;  /* DENSE: out; CONV2D: cout */
  int64_t bias_offset;   /* DENSE / CONV2D: offset of bias[out] in the flat parameter vector */
  int64_t kernel_offset; /* DENSE: offset of kernel[in,out]; CONV2D: of kernel[kh,kw,cin,cout] */
  int32_t kh, kw;        /* CONV2D: kernel height / width; INPUT: image height / width */
  int32_t stride, pad;   /* CONV2D: stride, zero rows / columns before the image; ZEROPAD: zeros added on each side */
} lip_layer_desc;

typedef enum { LIP_REGRESSOR = 0, LIP_CLASSIFIER = 1 } lip_model_type;

/* which output-space factor an operator applies (src/ggn.py:16-39,125-131) */
typedef enum {
  LIP_FACTOR_NONE = 0,  /* plain Jacobian (lla.py:153 predictive JVP) */
  LIP_FACTOR_SQRT = 1   /* L / L^T with L = diag(sqrt p) - p sqrt(p)^T (classifier) or exp(-logvar/2) (regressor) */
} lip_factor;

const char* lip_last_error(void);
int lip_version(void);
/* number of CUDA kernels this library has launched in this process (bench.py's gpu_launches evidence) */
int64_t lip_launch_count(void);
/* 1 if the current device is sm_100 (tcgen05 path usable), 0 otherwise, <0 on error */
int lip_device_is_sm100(void);

/* ---- model handle ------------------------------------------------------------------------------------
 * Replaces the `state` + `Z` closure capture of compute_ggn_vp / compute_W_vps (src/ggn.py:9-14,97-113). */
int lip_model_create(const lip_layer_desc* layers, int32_t n_layers, int32_t model_type,
                     int64_t num_params, lip_model** out);
int lip_model_destroy(lip_model* m);
int64_t lip_model_num_params(const lip_model* m);
int64_t lip_model_num_outputs(const lip_model* m);
int64_t lip_model_num_points(const lip_model* m);
/* 0: SIMT fp32 only; 1: tcgen05 3xTF32 for qualifying layers (default when the device is sm_100) */
int lip_model_set_tensor_path(lip_model* m, int32_t enable);
/* number of dense layers whose GEMMs run on the tcgen05 path for the current binding (0 = SIMT only) */
int lip_model_tensor_layers(const lip_model* m);
/* number of conv stages of a conv stage program (LeNet5, src/scalemodels.py:11-49) that run as ONE fused kernel per direction
 * (conv + activation mask + pool, no patch buffer: csrc/lip_cnn_fused.cu) for the current binding; 0 = im2col + GEMM */
int lip_model_fused_stages(const lip_model* m);

/* Bind weights theta[D] (flat, reference order) and points Z[M, in_features]; runs and caches the forward
 * pass (activations, activation derivatives, softmax p and sqrt p).  Replaces the per-call forward passes of
 * ggn.py:139-142 (three forwards per point per probe in the reference).  logvar is the regressor's
 * log-variance (ggn.py:112-113); ignored for classifiers.  Allocates the cache (not a hot call). */
int lip_model_bind(lip_model* m, const float* theta, const float* Z, int64_t M, float logvar,
                   lip_stream_t stream);
/* BatchNorm running statistics for programs with BATCHNORM ops (state.batch_stats, src/ggn.py:52): device array
 * [mean_0(c0), var_0(c0), mean_1(c1), var_1(c1), ...] in the order the BATCHNORM / RES_BATCHNORM ops appear; eps is
 * flax's 1e-5.  Must be called before lip_model_bind; the library keeps its own copy. */
int lip_model_set_bn_stats(lip_model* m, const float* stats, int64_t n, lip_stream_t stream);
/* Copies the cached model outputs f(theta, Z) [M,K] (logits / means) to out. */
int lip_model_outputs(lip_model* m, float* out, lip_stream_t stream);

/* scratch bytes the hot calls below need for B probes */
size_t lip_workspace_bytes(const lip_model* m, int64_t B);

/* ---- GGN-vector product ------------------------------------------------------------------------------
 * out[b,:] = recal * sum_i J_i^T H_i J_i V[b,:] + alpha * V[b,:]       V, out: [B, D] row-major.
 * Replaces ggn_vp (src/ggn.py:133-144) vmapped over probes (src/stochtrace.py:113-114) and
 * curvature_vp (src/lla.py:19-23).  recal = N/M (x exp(-logvar) for regressors, ggn.py:109-113). */
int lip_ggn_vp(lip_model* m, const float* V, float* out, int64_t B, float recal, float alpha,
               void* workspace, size_t workspace_bytes, lip_stream_t stream);
/* The same with explicit row strides (V: [B, ldv], out: [B, ldo], both >= D, floats) and flags.
 * LIP_PROBES_EXACT_TF32: the caller vouches that every entry of V is exactly representable in TF32 (10 mantissa bits) — the
 * reference's Rademacher +-1 probes (src/stochtrace.py:28, src/train_inducing.py:139) and one-hot blocks are.  When in addition the
 * rows are TMA-addressable (V 16-byte aligned, ldv % 4 == 0, kernel offsets and layer widths multiples of 4 floats) the tcgen05 JVP
 * GEMMs read the probe block IN PLACE and the TF32 split pass (12 % of a call) disappears.  A plain [B, D] block qualifies only when
 * D % 4 == 0; callers that own their probe buffers pad the rows (the MNIST MLP's D = 1,494,154 -> ldv = 1,494,156).  Without the flag,
 * or where a layer does not qualify, the call behaves exactly like lip_ggn_vp.  With the flag set on data that is NOT exactly TF32 the
 * low 13 mantissa bits of V are ignored (results lose fp32 accuracy; nothing else breaks). */
#define LIP_PROBES_EXACT_TF32 1
int lip_ggn_vp_ex(lip_model* m, const float* V, int64_t ldv, float* out, int64_t ldo, int64_t B, float recal, float alpha,
                  int32_t flags, void* workspace, size_t workspace_bytes, lip_stream_t stream);

/* out[b,i,:] = scale * F_i^T J_i V[b,:]      V: [B, D] -> out: [B, M, K]
 * factor=SQRT replaces WTfun (src/ggn.py:54-62,84-85); factor=NONE is the batched JVP of lla.py:153. */
int lip_wt_apply(lip_model* m, const float* V, float* out, int64_t B, float scale, int32_t factor,
                 void* workspace, size_t workspace_bytes, lip_stream_t stream);

/* out[b,:] = scale * sum_i J_i^T F_i U[b,i,:] + add_scale * add[b,:]   U: [B, M, K] -> out: [B, D]
 * Replaces Wfun (src/ggn.py:64-76,87-91) without materialising the [M,D] per-example gradients.
 * add may be NULL (then add_scale is ignored). */
int lip_w_apply(lip_model* m, const float* U, float* out, int64_t B, float scale, int32_t factor,
                const float* add, float add_scale,
                void* workspace, size_t workspace_bytes, lip_stream_t stream);

/* Dense Gram  G = scale^2 * W^T W  [d,d], d = M*K, symmetrised from the upper triangle as
 * src/ggn.py:227.  Replaces build_WTW (src/ggn.py:198-227): one batched W over one-hot blocks followed by
 * W^T instead of d/block sequential pairs.  block = one-hot columns pushed per pass. */
int lip_gram_wtw(lip_model* m, float* G, float scale, int64_t block,
                 void* workspace, size_t workspace_bytes, lip_stream_t stream);
size_t lip_gram_workspace_bytes(const lip_model* m, int64_t block);

/* Cross Gram  Gt = scale_x * scale_z * (W_x^T W_z)^T  [d_z, d_x] row-major (row j = W_x^T W_z e_j), d_x = M_x*K, d_z = M_z*K:
 * the same layer program bound at two point sets (mx at X, mz at Z; both handles share theta).  Replaces build_WTWz
 * (src/ggn.py:233-272), whose result [d_x, d_z] is the transpose view of Gt.  The workspace must satisfy
 * lip_gram_cross_workspace_bytes. */
int lip_gram_cross(lip_model* mx, lip_model* mz, float* Gt, float scale_x, float scale_z, int64_t block,
                   void* workspace, size_t workspace_bytes, lip_stream_t stream);
size_t lip_gram_cross_workspace_bytes(const lip_model* mx, const lip_model* mz, int64_t block);

/* ---- gradients with respect to the bound points Z ---------------------------------------------
 * The reference obtains dObjective/dZ by jax.value_and_grad THROUGH ggn_vp / Wfun / WTfun (src/train_inducing.py:195-232,
 * src/ggn.py:9-146).  These are the VJP-with-respect-to-Z rules a jax.custom_vjp around the entry points above needs
 * (the rule with respect to the vector argument is the operator itself / its adjoint).  With probes b < B, points i < M:
 *   LIP_ZGRAD_GGN : d/dZ sum_b X1[b]^T ( scale * sum_i J_i^T H_i J_i X2[b] )          X1 = cotangent [B,D], X2 = v [B,D]
 *   LIP_ZGRAD_WT  : d/dZ sum_b sum_i X2[b,i] . ( scale * L_i^T J_i X1[b] )            X1 = v [B,D], X2 = cotangent [B,M,K]
 *   LIP_ZGRAD_W   : d/dZ sum_b X1[b]^T ( scale * sum_i J_i^T L_i X2[b,i] )            X1 = cotangent [B,D], X2 = U [B,M,K]
 *   LIP_ZGRAD_JVP : d/dZ sum_b sum_i X2[b,i] . ( scale * J_i X1[b] )                  X1 = v [B,D], X2 = cotangent [B,M,K]
 * `scale` has the meaning of lip_ggn_vp's recal / lip_w(t)_apply's scale with factor = SQRT (NONE for LIP_ZGRAD_JVP).
 * out: [M, in] (per_probe = 0, summed over probes: what differentiating a vmapped closure yields) or [B, M, in] (per_probe = 1).
 * Built for dense programs (models M1 / M2: tcgen05 3xTF32 GEMMs for the layers on the tensor path when the activations are
 * tanh / relu, fp32 SIMT GEMMs otherwise) and for relu conv stage programs (model M3, LeNet5: Z = the input images);
 * residual programs (M4) return LIP_ERR_UNSUPPORTED. */
typedef enum { LIP_ZGRAD_GGN = 0, LIP_ZGRAD_WT = 1, LIP_ZGRAD_W = 2, LIP_ZGRAD_JVP = 3 } lip_zgrad_mode;
int lip_zgrad(lip_model* m, int32_t mode, const float* X1, const float* X2, float* out, int64_t B, float scale,
              int32_t per_probe, void* workspace, size_t workspace_bytes, lip_stream_t stream);
size_t lip_zgrad_workspace_bytes(const lip_model* m, int32_t mode, int64_t B);

/* ---- evaluation consumer ---------------------------------------------------------------------
 * Monte-Carlo softmax predictive from the S logit samples of predict_lla_scalable, as batch_nll computes it
 * (scale_experiments/evaluate.py:126-142):
 *   log_avg_prob[b] = logsumexp_s log_softmax(logits[s,b,:])[labels[b]] - log S     (NaN for labels outside [0, C))
 *   mean_probs[b,c] = mean_s softmax(logits[s,b,:])[c]
 * logits: [S, B, C] row-major; labels: int32 [B]; either output may be NULL (labels may be NULL without log_avg_prob). C <= 64. */
int lip_mc_softmax_predictive(const float* logits, const int32_t* labels, float* log_avg_prob, float* mean_probs,
                              int64_t S, int64_t B, int32_t C, lip_stream_t stream);

/* ---- vector stage (CG / Lanczos / GKL building blocks; all batched over B independent columns) --------
 * Vectors are [B, n] row-major.  Scalars stay on the device (no host readback). */

/* out[b] = sum_j x[b,j] * y[b,j]           (deterministic two-stage reduction; scratch >= lip_dot_scratch_bytes) */
size_t lip_dot_scratch_bytes(int64_t n, int64_t B);
int lip_dot(const float* x, const float* y, float* out, int64_t n, int64_t B, int64_t ldx, int64_t ldy,
            void* scratch, lip_stream_t stream);
/* y[b,:] = a[b] * x[b,:] + c[b] * y[b,:]   (a, c device arrays [B]; NULL means 1 / 0 resp.) */
int lip_axpby(const float* a, const float* x, const float* c, float* y, int64_t n, int64_t B,
              int64_t ldx, int64_t ldy, lip_stream_t stream);
/* y[b,:] = x[b,:] * (invert ? 1/s[b] : s[b]) */
int lip_scale(const float* s, int32_t invert, const float* x, float* y, int64_t n, int64_t B,
              int64_t ldx, int64_t ldy, lip_stream_t stream);

/* Rademacher probes in their compact wire format: row b of `bits` (ldbits bytes per row, numpy.packbits order: bit
 * 7 of byte 0 is element 0) -> out[b, j] = bit ? +1 : -1, j < n.  Host probes for Hutchinson / SLQ
 * (src/stochtrace.py:28, src/train_inducing.py:139) then cross PCIe at 1 bit per element instead of 32. */
int lip_unpack_rademacher(const uint8_t* bits, int64_t ldbits, float* out, int64_t n, int64_t B, lip_stream_t stream);
/* the same into rows of stride ldo >= n floats (padded probe rows for lip_ggn_vp_ex) */
int lip_unpack_rademacher_ld(const uint8_t* bits, int64_t ldbits, float* out, int64_t ldo, int64_t n, int64_t B, lip_stream_t stream);

/* One CG iteration's vector work for jax.scipy.sparse.linalg.cg semantics (call sites
 * src/stochtrace.py:146,192; src/sample.py:71): given Ap = A p,
 *   a = gamma/(p.Ap); x += a p; r -= a Ap; gamma' = r.r; p = r + (gamma'/gamma) p; gamma = gamma'
 * for every column b with active[b] != 0; active[b] is then recomputed as (gamma'[b] > thresh[b]).
 * iters[b] is incremented for active columns.  state arrays gamma, thresh: [B] fp32; active, iters: [B] int32. */
int lip_cg_step(float* x, float* r, float* p, const float* Ap, float* gamma, const float* thresh,
                int32_t* active, int32_t* iters, int64_t n, int64_t B, void* scratch, lip_stream_t stream);
/* r = p = b, x = 0, gamma = b.b, thresh = max(tol^2 b.b, atol^2), active = gamma > thresh, iters = 0 */
int lip_cg_init(const float* b, float* x, float* r, float* p, float* gamma, float* thresh,
                int32_t* active, int32_t* iters, float tol, float atol, int64_t n, int64_t B,
                void* scratch, lip_stream_t stream);

/* Full re-orthogonalisation against the first kk rows of a basis Q[B, kmax, ldq] (row stride ldq >= n,
 * batch stride kmax*ldq; pad ldq to a multiple of 4 for 128-bit loads):
 *   h = Q w;  w -= Q^T h;  [if passes == 2:  w -= Q^T (Q w)  with the 2nd coefficients discarded]
 * w: [B, ldw].  h_out[B, kmax] (optional) receives the first-pass coefficients (entries >= kk untouched).
 * norm_out[B] (optional) receives |w| after the update.  This is the Gram-Schmidt of matfree's
 * decomp.tridiag_sym / decomp.bidiag with reortho="full" (call sites src/sample.py:114,
 * src/train_inducing.py:156).  scratch >= lip_reorth_scratch_bytes. */
size_t lip_reorth_scratch_bytes(int64_t n, int64_t B, int64_t kmax);
int lip_reorth(const float* Q, int64_t ldq, int64_t kk, int64_t kmax, float* w, int64_t ldw, float* h_out,
               float* norm_out, int32_t passes, int64_t n, int64_t B, void* scratch, lip_stream_t stream);
/* out[b,:] = sum_{j<kk} c[b,j] * Q[b,j,:]      (c: [B, ldc], out: [B, ldo]) */
int lip_basis_combine(const float* Q, int64_t ldq, int64_t kk, int64_t kmax, const float* c, int64_t ldc,
                      float* out, int64_t ldo, int64_t n, int64_t B, lip_stream_t stream);

/* ---- small dense stage ---------------------------------------------------------------------------------
 * Symmetric tridiagonal eigen-decomposition + matrix function, one thread block per problem, float64 inside.
 * diag[B,k], off[B,k-1] (fp32 in).  fn: 0 log, 1 inverse sqrt, 2 inverse, 3 identity.
 * clip_min < 0 disables the clip; the reference's patched eigh clips eigenvalues to >= 1.0
 * (src/matfree_monkeypatch.py:19).
 *   quad_out[B]   (optional) = e1^T f(T) e1                       (matfree funm.integrand_funm_sym)
 *   fe1_out[B,k]  (optional) = f(T) e1                            (matfree funm.funm_lanczos_sym)
 *   eig_out[B,k]  (optional) = eigenvalues (ascending not guaranteed)
 * Replaces dense_funm_sym_eigh (src/matfree_monkeypatch.py:8-22) and the dense SVD of matfree's
 * dense_funm_product_svd when fed T = B^T B (see lip_bidiag_to_tridiag). */
size_t lip_tridiag_scratch_bytes(int64_t k, int64_t B, int32_t want_vectors);
int lip_tridiag_funm(const float* diag, const float* off, int64_t k, int64_t B, int32_t fn, float clip_min,
                     float* quad_out, float* fe1_out, float* eig_out, void* scratch, lip_stream_t stream);
/* The same with parameterised functions.  fn 4 = LIP_FN_SAMPLER, params = {alpha, beta, tau} (a HOST array of 3 floats):
 *   psi(th) = (clip(th, clip_min)^{-1/2} - alpha^{-1/2}) / lam,  lam = (th - alpha)/beta;  psi = 0 where lam <= tau * lam_max.
 * On the Lanczos matrix of alpha I + beta W^T W (src/sample.py:113-125) this is the whole output-space part of the posterior
 * sampler: A^{-1/2} v = alpha^{-1/2} v + W psi(W^T W) W^T v — the two solves with the singular Gram of src/sample.py:81,135
 * (jax.scipy.linalg.solve) become part of the matrix function, with the pseudo-inverse cut at the fp32 noise level tau. */
int lip_tridiag_funm_p(const float* diag, const float* off, int64_t k, int64_t B, int32_t fn, float clip_min, const float* params,
                       float* quad_out, float* fe1_out, float* eig_out, void* scratch, lip_stream_t stream);
/* T = Bd^T Bd for upper-bidiagonal Bd = diag(alphas) + superdiag(betas[1:]):  tdiag[i] = a_i^2 + b_i^2
 * (b_0 := 0), toff[i] = a_i * b_{i+1}.  alphas, betas: [B,k]. */
int lip_bidiag_to_tridiag(const float* alphas, const float* betas, float* tdiag, float* toff, int64_t k,
                          int64_t B, lip_stream_t stream);

/* ---- composite Krylov entry points -------------------------------------------------------------------------------
 * ONE call enqueues a whole recurrence on the caller's stream: no host synchronisation (lip_cg_solve: optional), no
 * allocation (caller workspace, lip_krylov_workspace_bytes), every scalar on the device, CUDA-graph capturable for the
 * built-in operator kinds.  These are what an XLA custom call would bind for matfree.decomp.tridiag_sym / decomp.bidiag /
 * funm.* and jax.scipy.sparse.linalg.cg at the reference's call sites (src/sample.py:71,113-115; src/train_inducing.py:156-163;
 * src/stochtrace.py:146,192; src/matfree_monkeypatch.py:25-41). */

/* A linear operator for the Krylov routines. */
typedef enum {
  LIP_LINOP_GGN = 0,        /* v[D] -> scale * sum_i J_i^T H_i J_i v + alpha v   (curvature_vp, src/lla.py:19-23; symmetric) */
  LIP_LINOP_GKL = 1,        /* v[D] -> [sqrt(alpha) v ; scale * W^T v] in R^{D+d}  (bidiag_target, src/train_inducing.py:166-169);
                               transpose u -> sqrt(alpha) u[:D] + scale * W u[D:]  (what jax.vjp derives inside matfree) */
  LIP_LINOP_DENSE_SYM = 2,  /* u[n] -> alpha u + beta * G u, G [n,n] symmetric, row-major (inner_fun_flat, src/sample.py:120-125) */
  LIP_LINOP_CALLBACK = 3    /* a caller function (any other composition, e.g. S_X S_Z^{-1} of src/train_inducing.py:134-135) */
} lip_linop_kind;

/* CALLBACK protocol: the library copies the operator's input to cb_in ([B, n] contiguous; transpose: cb_in_t [B, n_out]), calls
 * fn(ctx, transpose, B, stream), and reads the result from cb_out ([B, n_out]; transpose: cb_out_t [B, n]).  fn enqueues its
 * work on `stream` and returns 0, or non-zero to abort the recurrence.  symmetric != 0: fn is never called with transpose = 1. */
typedef int (*lip_matvec_fn)(void* ctx, int32_t transpose, int64_t B, lip_stream_t stream);

typedef struct {
  int32_t kind;          /* lip_linop_kind */
  int32_t symmetric;     /* CALLBACK: 1 if A = A^T */
  lip_model* model;      /* GGN, GKL: a bound model */
  float scale;           /* GGN: recal = N/M (x exp(-logvar));  GKL: the scale of W (sqrt(N/M); 1 at the reference's call site) */
  float alpha;           /* GGN: + alpha v;  GKL: alpha (its square root is taken inside);  DENSE_SYM: alpha */
  float beta;            /* DENSE_SYM */
  const float* dense;    /* DENSE_SYM: G, device */
  int64_t n;             /* DENSE_SYM, CALLBACK: input dimension (model kinds: D) */
  int64_t n_out;         /* CALLBACK: output dimension */
  lip_matvec_fn fn;      /* CALLBACK */
  void* ctx;
  float *cb_in, *cb_out, *cb_in_t, *cb_out_t;   /* CALLBACK: device buffers, see above (the _t pair may be NULL when symmetric) */
} lip_linop;

typedef enum {
  LIP_KRYLOV_LANCZOS = 0, LIP_KRYLOV_GKL = 1, LIP_KRYLOV_SLQ_LANCZOS = 2, LIP_KRYLOV_SLQ_GKL = 3, LIP_KRYLOV_FUNM = 4,
  LIP_KRYLOV_CG = 5, LIP_KRYLOV_HUTCHPP = 6, LIP_KRYLOV_APPLY = 7
} lip_krylov_routine;
/* workspace bytes of the routine below for depth k (CG, APPLY: ignored; HUTCHPP: k = s1, B = s2) and B columns */
size_t lip_krylov_workspace_bytes(const lip_linop* op, int32_t routine, int64_t k, int64_t B);

/* out = A in (transpose = 0: in [B, n] -> out [B, n_out]) or A^T in (transpose = 1), both contiguous: the operator on its own. */
int lip_linop_apply(const lip_linop* op, const float* in, float* out, int64_t B, int32_t transpose, void* workspace,
                    size_t workspace_bytes, lip_stream_t stream);

/* matfree.decomp.tridiag_sym(k) with full re-orthogonalisation (passes = 2: CGS2, what the Python mirror used; 1: CGS), batched:
 * q_0 = v0/|v0|; for i < k: w = A q_i; h = Q^T w (first pass = Arnoldi column); w -= Q h [twice]; q_{i+1} = w/|w|.
 * diag[B,k], off[B,k-1] receive T = (H + H^T)/2 on its three diagonals; norm0[B] (optional) = |v0|.
 * Q [B, k, ldq] is the caller-owned basis (16-byte aligned, ldq % 4 == 0, rows zero-padded).  v0: [B, ldv0]. */
int lip_lanczos_tridiag(const lip_linop* op, const float* v0, int64_t ldv0, int64_t k, int64_t B, int32_t passes, float* Q,
                        int64_t ldq, float* diag, float* off, float* norm0, void* workspace, size_t workspace_bytes,
                        lip_stream_t stream);

/* matfree.decomp.bidiag(k): Golub-Kahan-Lanczos with full re-orthogonalisation of both bases, batched.  alphas[B,k] (diagonal),
 * betas[B,k] (betas[:,0] = 0; betas[:,i] = super-diagonal entry (i-1, i)).  Us [B,k,ldu], Vs [B,k,ldv] caller-owned bases. */
int lip_gkl_bidiag(const lip_linop* op, const float* v0, int64_t ldv0, int64_t k, int64_t B, float* Us, int64_t ldu, float* Vs,
                   int64_t ldv, float* alphas, float* betas, float* norm0, void* workspace, size_t workspace_bytes,
                   lip_stream_t stream);

/* Stochastic-Lanczos-quadrature integrand for every probe row: quad_out[b] = |v_b|^2 e1^T f(T_b) e1.
 * form LIP_SLQ_LANCZOS: T from lip_lanczos_tridiag (matfree funm.integrand_funm_sym; with fn = log, clip_min = 1 the patched
 *   integrand_funm_sym_logdet of src/matfree_monkeypatch.py:25-41);
 * form LIP_SLQ_GKL: T = B^T B from lip_gkl_bidiag (funm.integrand_funm_product_logdet, src/train_inducing.py:156-157).
 * fn / clip_min as lip_tridiag_funm.  The mean over probes (matfree stochtrace.estimator) is the caller's (and, sharded over
 * GPUs, one all-reduce of B floats). */
typedef enum { LIP_SLQ_LANCZOS = 0, LIP_SLQ_GKL = 1 } lip_slq_form;
int lip_slq_quadrature(const lip_linop* op, const float* probes, int64_t ldp, int64_t k, int64_t B, int32_t form, int32_t fn,
                       float clip_min, float* quad_out, void* workspace, size_t workspace_bytes, lip_stream_t stream);

/* ---- multi-GPU: the same recurrence with its Krylov bases sharded over the ranks of a communicator ----------------------
 * lip_comm is a NCCL communicator the library owns (NCCL is resolved at run time from the process: no link-time dependency).  Rank 0
 * of a group obtains 128 id bytes with lip_comm_unique_id, the caller ships them to the other ranks (torch.distributed in the Python
 * mirror), every rank calls lip_comm_create with the CUDA device it will use current. */
typedef struct lip_comm lip_comm;
int lip_comm_unique_id(void* id128);
int lip_comm_create(const void* id128, int32_t world, int32_t rank, lip_comm** out);
int lip_comm_destroy(lip_comm* comm);
int lip_comm_world(const lip_comm* comm);
int lip_comm_rank(const lip_comm* comm);
int lip_comm_allreduce_sum(lip_comm* comm, float* buf, int64_t count, lip_stream_t stream);
/* lip_slq_quadrature with the basis rows cut column-wise over comm's ranks: rank r owns n/S columns of every Krylov vector, so the
 * O(k^2 n) re-orthogonalisation traffic that bounds the logdet is divided by S; the operator is applied replicated on the
 * all-gathered vector (its latency does not shrink with the batch); norms and re-orthogonalisation coefficients are all-reduced
 * (B resp. B x k floats per exchange).  Every rank passes the same FULL probes and receives the same quad_out.  comm == NULL or
 * a single rank: lip_slq_quadrature.  Model operator kinds only.  Workspace: lip_slq_workspace_bytes(op, form, k, B, world). */
int lip_slq_quadrature_sharded(const lip_linop* op, lip_comm* comm, const float* probes, int64_t ldp, int64_t k, int64_t B, int32_t form,
                               int32_t fn, float clip_min, float* quad_out, void* workspace, size_t workspace_bytes,
                               lip_stream_t stream);
size_t lip_slq_workspace_bytes(const lip_linop* op, int32_t form, int64_t k, int64_t B, int32_t world);

/* matfree.funm.funm_lanczos_sym: out[b,:] = |v_b| Q_b f(T_b) e1 ~= f(A) v_b   (src/sample.py:113-115: f = 1/sqrt, clip_min = 1).
 * fn_params: host array for parameterised functions (lip_tridiag_funm_p), else NULL. */
int lip_funm_lanczos(const lip_linop* op, const float* v, int64_t ldv, int64_t k, int64_t B, int32_t fn, float clip_min,
                     const float* fn_params, float* out, int64_t ldo, void* workspace, size_t workspace_bytes, lip_stream_t stream);

/* hutchpp_v2 (src/stochtrace.py:118-135):  tr X ~= tr(Q^T X Q) + tr(G_perp X G_perp^T) / s2,  S = probes[:s1], G = probes[s1:s1+s2],
 * Q = orth(X S^T), G_perp = G - (G Q) Q^T.  probes: [s1 + s2, ldp] rows.  out: one float (device).
 * jnp.linalg.qr of the [n, s1] block is a shifted CholeskyQR with three passes (float64 Gram by a tall-skinny reduction kernel, one-CTA
 * Cholesky + triangular inverse, Q <- L^{-1} Y as a GEMM): the estimate depends on span(Q) only, so any orthonormal basis is the
 * reference's.  info (optional, int32 device): non-zero when a Cholesky pivot was not positive (cond(X S^T) beyond ~1e7).
 * Workspace: lip_krylov_workspace_bytes(op, LIP_KRYLOV_HUTCHPP, s1, s2). */
int lip_hutchpp_v2(const lip_linop* op, const float* probes, int64_t ldp, int64_t s1, int64_t s2, float* out, int32_t* info,
                   void* workspace, size_t workspace_bytes, lip_stream_t stream);

/* jax.scipy.sparse.linalg.cg(A, b): x0 = 0, stop when r.r <= max(tol^2 b.b, atol^2) or after maxiter iterations (< 0: 10 n).
 * b, x: [B, n] contiguous.  check_every > 0: the host polls a pinned flag every check_every iterations and stops enqueueing once
 * every column has converged (it synchronises with the stream before returning); check_every = 0: exactly maxiter masked
 * iterations are enqueued and the call never waits (graph-capturable).  iters_out: int32 [B] device, optional. */
int lip_cg_solve(const lip_linop* op, const float* b, float* x, int64_t B, float tol, float atol, int64_t maxiter,
                 int32_t check_every, int32_t* iters_out, void* workspace, size_t workspace_bytes, lip_stream_t stream);

/* ---- self test of the tensor-core GEMM (3xTF32 tcgen05) against the SIMT fp32 GEMM; returns max rel err
 * through *max_rel_err.  variant: 0 = JVP-type (A K-major, B N-major), 1 = weight-grad type (both MN-major),
 * 2 = delta-backprop type (both K-major). */
int lip_selftest_tc_gemm(int32_t variant, int64_t Mrows, int64_t N, int64_t K, int64_t batch,
                         float* max_rel_err, lip_stream_t stream);

/* Microbenchmark of the tensor-core GEMM (zero operands, same variants as the self test): average ms per launch
 * over `iters` launches.  dbg knobs (diagnosis only; results are then meaningless): 1 = skip the epilogue's global
 * traffic, 2 = issue only the hi*hi MMA, 4 = skip the TMA loads.  two_cta: 0 = 1-CTA kernel, 1 = cta_group::2. */
int lip_bench_tc_gemm(int32_t variant, int64_t Mrows, int64_t N, int64_t K, int64_t batch, int32_t iters,
                      int32_t dbg, int32_t two_cta, float* ms_per_iter, lip_stream_t stream);

/* Self test / microbenchmark of the tcgen05 implicit-GEMM convolutions (lip_conv_tc.cu) against the fp32 SIMT implicit GEMM
 * on random data, for a SAME ksz x ksz conv (stride 1 or 2) on `imgs` images [H, W, cin] -> cout and `batch` probes.
 * role: 0 = JVP (shared image x per-probe kernels + per-probe image x shared kernel), 1 = per-probe kernel gradient,
 * 2 = delta back-propagation (transposed conv), 3 = the first JVP term alone with the probes folded into the tile width.
 * *rel_err = relative L2 error; with iters > 0, *ms_tc / *ms_simt = mean
 * time of one tensor-core / one SIMT call. */
int lip_selftest_conv_tc(int32_t role, int64_t imgs, int32_t H, int32_t W, int32_t cin, int32_t cout, int32_t ksz,
                         int32_t stride, int64_t batch, int32_t iters, float* rel_err, float* ms_tc, float* ms_simt, lip_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* LIP_B200_H_ */
