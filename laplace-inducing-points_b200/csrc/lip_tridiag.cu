// lip_tridiag.cu — on-device symmetric tridiagonal eigensolver + matrix function, one CTA per problem.
//
// Reference semantics: dense_funm_sym_eigh (src/matfree_monkeypatch.py:8-22: eigh, clip(eigvals, min=1.0),
// V f(L) V^T) as consumed by matfree's funm_lanczos_sym (f(T) e1, src/sample.py:113-115) and
// integrand_funm_sym (e1^T f(T) e1, src/matfree_monkeypatch.py:25-41); and matfree's
// dense_funm_product_svd on the GKL bidiagonal (src/train_inducing.py:156-157) through T = B^T B.
//
// Algorithm: implicit-shift QL (EISPACK tql2 recurrences) in float64.  Thread 0 runs the scalar recurrence
// of one QL sweep and records its Givens rotations in shared memory; then all threads apply the sweep's
// rotations to their own rows of the eigenvector matrix (stored transposed so that accesses are coalesced).
// For quadrature only the first row is carried (O(k^2) total work).
#include <math.h>

#include "lip_common.cuh"

using namespace lip;

namespace {

__device__ __forceinline__ double apply_fn(int fn, double x) {
  switch (fn) {
    case 0: return log(x);
    case 1: return 1.0 / sqrt(x);
    case 2: return 1.0 / x;
    default: return x;
  }
}

// fn 4 (LIP_FN_SAMPLER): the posterior sampler's matrix function of the Gram G = W^T W, evaluated on the Ritz values th of
// alpha I + beta G (the operator src/sample.py:120-125 hands to Lanczos):
//     psi(th) = ( clip(th, clip_min)^{-1/2} - alpha^{-1/2} ) / lam,   lam = (th - alpha) / beta   (the Ritz value of G),
// and psi = 0 where lam <= tau * lam_max (the pseudo-inverse of the rank-deficient Gram: fp32 cannot tell those directions from
// null(W)).  With y = W^T v it gives  A^{-1/2} v = alpha^{-1/2} v + W psi(G) y  — the formula of src/sample.py:117-143
// (Higham et al., thm 1.2) with both singular solves (G^+ applied to f(..) y and to y) folded into the matrix function.
__device__ __forceinline__ double sampler_fn(double th, double th_max, float clip_min, double alpha, double beta, double tau) {
  const double lam = (th - alpha) / beta, lam_max = (th_max - alpha) / beta;
  const bool clipped = clip_min >= 0.f && alpha < (double)clip_min && th < (double)clip_min;
  if (!clipped) {
    // ((alpha + beta lam)^{-1/2} - alpha^{-1/2}) / lam without the cancellation: finite and smooth through lam = 0, so the noise-level
    // Ritz values of the rank-deficient Gram (either sign) need no cut here
    const double t = th > 1e-30 ? th : 1e-30;
    const double st = sqrt(t), sa = sqrt(alpha);
    return -beta / (st * sa * (st + sa));
  }
  // clipped branch (alpha < clip_min): (clip_min^{-1/2} - alpha^{-1/2}) / lam is singular at lam = 0 — the reference's formula jumps
  // from clip_min^{-1/2} on range(W) to alpha^{-1/2} on null(W); directions below the cut count as null(W)
  if (!(lam > tau * lam_max)) return 0.0;
  return (1.0 / sqrt((double)clip_min) - 1.0 / sqrt(alpha)) / lam;
}

// scratch per problem (doubles): d[n] e[n] cs[n] sn[n] Zt[nrows*n]
__global__ void tridiag_funm_kernel(const float* __restrict__ diag, const float* __restrict__ off, int n, int fn,
                                    float clip_min, float p0, float p1, float p2, float* __restrict__ quad_out,
                                    float* __restrict__ fe1_out, float* __restrict__ eig_out, double* __restrict__ scratch,
                                    int nrows, int use_smem) {
  extern __shared__ double smd[];
  const int b = blockIdx.x;
  const size_t per = (size_t)4 * n + (size_t)nrows * n;
  double* base = scratch + (size_t)b * per;
  double* d = use_smem ? smd : base;
  double* e = d + n;
  double* cs = e + n;
  double* sn = cs + n;
  double* Zt = base + (size_t)4 * n;  // Zt[i*nrows + k] = Z[k][i]
  __shared__ int sh_lo, sh_hi, sh_done, sh_fail;

  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    d[i] = (double)diag[(size_t)b * n + i];
    e[i] = (i < n - 1) ? (double)off[(size_t)b * (n - 1) + i] : 0.0;
  }
  for (size_t idx = threadIdx.x; idx < (size_t)nrows * n; idx += blockDim.x) {
    size_t i = idx / nrows, k = idx % nrows;
    Zt[idx] = (i == k) ? 1.0 : 0.0;
  }
  if (threadIdx.x == 0) { sh_done = 0; sh_fail = 0; }
  __syncthreads();

  int l = 0, iter = 0;  // only meaningful on thread 0
  while (true) {
    if (threadIdx.x == 0) {
      int lo = 0, hi = 0;  // rotations recorded for i in [lo, hi)
      bool produced = false;
      while (!produced) {
        if (l >= n) { sh_done = 1; break; }
        int m = l;
        for (; m < n - 1; ++m) {
          double dd = fabs(d[m]) + fabs(d[m + 1]);
          if (fabs(e[m]) + dd == dd) break;
        }
        if (m == l) { ++l; iter = 0; continue; }
        if (++iter > 200) { sh_fail = 1; sh_done = 1; break; }
        double g = (d[l + 1] - d[l]) / (2.0 * e[l]);
        double r = hypot(g, 1.0);
        g = d[m] - d[l] + e[l] / (g + copysign(r, g));
        double s = 1.0, c = 1.0, p = 0.0;
        int i = m - 1;
        bool early = false;
        for (; i >= l; --i) {
          double f = s * e[i], bb = c * e[i];
          r = hypot(f, g);
          e[i + 1] = r;
          if (r == 0.0) {
            d[i + 1] -= p;
            e[m] = 0.0;
            early = true;
            break;
          }
          s = f / r;
          c = g / r;
          g = d[i + 1] - p;
          r = (d[i] - g) * s + 2.0 * c * bb;
          p = s * r;
          d[i + 1] = g + p;
          g = c * r - bb;
          cs[i] = c;
          sn[i] = s;
        }
        if (!early) {
          d[l] -= p;
          e[l] = g;
          e[m] = 0.0;
        }
        lo = early ? i + 1 : l;
        hi = m;
        if (hi > lo) produced = true;
      }
      sh_lo = lo;
      sh_hi = hi;
    }
    __syncthreads();
    if (sh_done) break;
    const int lo = sh_lo, hi = sh_hi;
    for (int k = threadIdx.x; k < nrows; k += blockDim.x) {
      double zh = Zt[(size_t)hi * nrows + k];
      for (int i = hi - 1; i >= lo; --i) {
        const double zl = Zt[(size_t)i * nrows + k];
        const double c = cs[i], s = sn[i];
        Zt[(size_t)(i + 1) * nrows + k] = s * zl + c * zh;
        zh = c * zl - s * zh;
      }
      Zt[(size_t)lo * nrows + k] = zh;
    }
    __syncthreads();
  }
  __syncthreads();
  const bool fail = sh_fail != 0;
  __shared__ double sh_max;
  if (fn == 4) {
    if (threadIdx.x == 0) {
      double mx = d[0];
      for (int m2 = 1; m2 < n; ++m2) mx = fmax(mx, d[m2]);
      sh_max = mx;
    }
    __syncthreads();
  }
  // g[m] = f(clip(lambda_m)) * Z[0][m]  (stored in cs);  quad = sum_m g[m] * Z[0][m]
  for (int m2 = threadIdx.x; m2 < n; m2 += blockDim.x) {
    double lam = d[m2];
    if (eig_out) eig_out[(size_t)b * n + m2] = fail ? nanf("") : (float)lam;
    double fv;
    if (fn == 4) {
      fv = sampler_fn(lam, sh_max, clip_min, (double)p0, (double)p1, (double)p2);
    } else {
      if (clip_min >= 0.f && lam < (double)clip_min) lam = (double)clip_min;
      fv = apply_fn(fn, lam);
    }
    cs[m2] = fv * Zt[(size_t)m2 * nrows + 0];
  }
  __syncthreads();
  if (quad_out && threadIdx.x == 0) {
    double q = 0.0;
    for (int m2 = 0; m2 < n; ++m2) q += cs[m2] * Zt[(size_t)m2 * nrows + 0];
    quad_out[b] = fail ? nanf("") : (float)q;
  }
  if (fe1_out && nrows == n) {
    for (int j = threadIdx.x; j < n; j += blockDim.x) {
      double acc = 0.0;
      for (int m2 = 0; m2 < n; ++m2) acc += Zt[(size_t)m2 * nrows + j] * cs[m2];
      fe1_out[(size_t)b * n + j] = fail ? nanf("") : (float)acc;
    }
  }
}

__global__ void bidiag_to_tridiag_kernel(const float* __restrict__ al, const float* __restrict__ be,
                                         float* __restrict__ td, float* __restrict__ to, int k, int B) {
  int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= B * k) return;
  int b = idx / k, i = idx % k;
  double a = al[(size_t)b * k + i];
  double bi = (i > 0) ? (double)be[(size_t)b * k + i] : 0.0;
  td[(size_t)b * k + i] = (float)(a * a + bi * bi);
  if (i < k - 1) to[(size_t)b * (k - 1) + i] = (float)(a * (double)be[(size_t)b * k + i + 1]);
}

}  // namespace

extern "C" {

size_t lip_tridiag_scratch_bytes(int64_t k, int64_t B, int32_t want_vectors) {
  size_t nrows = want_vectors ? (size_t)k : 1;
  return sizeof(double) * (size_t)B * (4 * (size_t)k + nrows * (size_t)k) + 256;
}

int lip_tridiag_funm(const float* diag, const float* off, int64_t k, int64_t B, int32_t fn, float clip_min,
                     float* quad_out, float* fe1_out, float* eig_out, void* scratch, lip_stream_t stream) {
  LIP_REQUIRE(fn >= 0 && fn <= 3, "lip_tridiag_funm: unknown function %d", fn);
  return lip_tridiag_funm_p(diag, off, k, B, fn, clip_min, nullptr, quad_out, fe1_out, eig_out, scratch, stream);
}

int lip_tridiag_funm_p(const float* diag, const float* off, int64_t k, int64_t B, int32_t fn, float clip_min, const float* params,
                       float* quad_out, float* fe1_out, float* eig_out, void* scratch, lip_stream_t stream) {
  LIP_REQUIRE(diag && (off || k == 1) && scratch && k > 0 && B > 0, "lip_tridiag_funm: bad argument");
  LIP_REQUIRE(fn >= 0 && fn <= 4, "lip_tridiag_funm: unknown function %d", fn);
  LIP_REQUIRE(fn != 4 || (params && params[0] > 0.f && params[1] > 0.f && params[2] >= 0.f),
              "lip_tridiag_funm: LIP_FN_SAMPLER needs params = {alpha > 0, beta > 0, tau >= 0} (host array)");
  const float p0 = params ? params[0] : 0.f, p1 = params ? params[1] : 0.f, p2 = params ? params[2] : 0.f;
  LIP_REQUIRE(quad_out || fe1_out || eig_out, "lip_tridiag_funm: no output requested");
  cudaStream_t st = (cudaStream_t)stream;
  const int n = (int)k;
  const int nrows = fe1_out ? n : 1;
  int threads = 32;
  if (nrows > 1) {
    threads = (n + 31) / 32 * 32;
    if (threads > 1024) threads = 1024;
  }
  size_t smem = sizeof(double) * 4 * (size_t)n;
  int use_smem = smem <= 96 * 1024;
  if (use_smem && smem > 48 * 1024) {
    LIP_CHECK_CUDA(cudaFuncSetAttribute(tridiag_funm_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024));
  }
  double* sc = (double*)align_up((uintptr_t)scratch, 16);
  tridiag_funm_kernel<<<(unsigned)B, threads, use_smem ? smem : 0, st>>>(diag, off, n, fn, clip_min, p0, p1, p2, quad_out,
                                                                         fe1_out, eig_out, sc, nrows, use_smem);
  LIP_LAUNCH_CHECK();
  return LIP_OK;
}

int lip_bidiag_to_tridiag(const float* alphas, const float* betas, float* tdiag, float* toff, int64_t k, int64_t B,
                          lip_stream_t stream) {
  LIP_REQUIRE(alphas && betas && tdiag && (toff || k == 1) && k > 0 && B > 0, "lip_bidiag_to_tridiag: bad argument");
  bidiag_to_tridiag_kernel<<<(unsigned)ceil_div(B * k, 256), 256, 0, (cudaStream_t)stream>>>(alphas, betas, tdiag, toff,
                                                                                            (int)k, (int)B);
  LIP_LAUNCH_CHECK();
  return LIP_OK;
}

}  // extern "C"
