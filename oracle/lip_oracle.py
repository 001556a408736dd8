"""CPU oracle for the matrix-free linearized-Laplace hot path.  TEST INFRASTRUCTURE ONLY.

Each function cites the reference file:line it restates (paths relative to /root/reference).
Arithmetic is float64 (torch CPU for autodiff, numpy elsewhere) unless a dtype is passed
(the fp32 structure-faithful variants are what bench.py times as the CPU baseline).

PARITY PINNING
  * pinned: G1 (tests/fixtures.py:24 linear model GGN = e^{-logvar}[[14.46,3.6],[3.6,4]]),
    G2 (tests/test_sample.py:334-355 Lanczos-20 inverse sqrt on diag(1..100)/100, rtol 1e-1),
    G3 (tests/fixtures.py:201-209 traces 6 / 3894, hutchpp_v2 exact for s1>=n, tests/test_stochtrace.py:90-97),
    G4 identities (vmap(ggn_vp)(I)==dense GGN, W(W^T(I))==GGN, L L^T==diag(p)-pp^T, null-space projector).
    See tests/test_oracle_golden.py.
  * PARITY UNPINNED (against the third-party SOURCES): matfree (requirements.txt:5, no version pin, source absent from
    /root/reference and not installed) and jax (absent).  `tridiag_sym`, `bidiag`, `funm_lanczos_sym`, `integrand_funm_*`,
    `estimator` and `cg` below restate the published algorithms of matfree>=0.1 / jax.scipy.sparse.linalg.cg from memory.
    No reference test touches decomp.bidiag + integrand_funm_product_logdet or the clip(min=1.0) of matfree_monkeypatch.py:19.
  * What IS pinned for them (round 2): the mathematics, per probe, against an INDEPENDENT dense ground truth on the reference's own
    operators — explicit Jacobians -> W -> float64 eigen-decomposition of alpha I + beta W W^T:
      tests/test_oracle_golden.py::test_slq_integrands_on_the_model_operators_equal_dense_quadratic_forms   (1e-9, small model)
      tests/golden/configs_v2.npz `c3a_dense_*` (C3a at its own size, M = 512, k = 409; checked here to 1e-13 while generating and
      against the CUDA path in tests/test_gpu_config_parity.py), and G2.
    What stays unpinned is matfree's / JAX's rounding ORDER — irrelevant at the 1e-4 tolerance wherever the quadrature is converged,
    and the reason unconverged clipped-Lanczos quadratures are compared at 5e-4 (see the C3b test).
  * `cg(..., dtype=np.float32)` runs the same recurrence in the reference's production precision: the yardstick for operators whose
    condition number exceeds 1 / eps_fp32 (C2: 2e6), where no fp32 CG — JAX's included — reaches tol = 1e-5.
"""
from __future__ import annotations

import math
from typing import Callable, Optional

import numpy as np
import torch

from .models import OracleState, flatten_nn_params

F64 = torch.float64


# ======================================================================================
# helpers
# ======================================================================================
def _t(x, dtype=F64):
    return torch.as_tensor(np.asarray(x), dtype=dtype)


def softmax_np(f):
    f = f - f.max(axis=-1, keepdims=True)
    e = np.exp(f)
    return e / e.sum(axis=-1, keepdims=True)


def _model_fn(state: OracleState, Z, dtype=F64):
    """theta -> f(theta, Z) [M, K]; per ggn.py:42-52,115-123 (BN eval mode, return_logvar=False)."""
    Zt = _t(Z, dtype)

    def f(theta):
        return state.f_theta(theta, Zt)

    return f


def model_outputs(state: OracleState, Z) -> np.ndarray:
    theta, _ = state.flat()
    with torch.no_grad():
        return _model_fn(state, Z)(_t(theta)).numpy()


def jvp_outputs(state: OracleState, X, w) -> np.ndarray:
    """lla.py:153: jax.jvp(lambda p: model(p, X), (theta,), (w,))[1]  ->  J_X w, shape [n, K]."""
    theta, _ = state.flat()
    return torch.func.jvp(_model_fn(state, X), (_t(theta),), (_t(w),))[1].detach().numpy()


def jacobians(state: OracleState, Z) -> np.ndarray:
    """Explicit per-point Jacobians J[i] = d f(z_i;theta)/d theta, shape [M, K, D] (ground truth)."""
    theta, _ = state.flat()
    th = _t(theta)
    Zt = _t(Z)
    rows = []
    for i in range(Zt.shape[0]):
        zi = Zt[i:i + 1]
        J = torch.func.jacrev(lambda p: state.f_theta(p, zi)[0])(th)
        rows.append(J.reshape(-1, th.numel()).numpy())
    return np.stack(rows)


# ======================================================================================
# src/ggn.py
# ======================================================================================
def _H_action(model_type, f_out, u):
    """ggn.py:125-131 : (diag(p) - p p^T) u  for classifiers, identity for regressors."""
    if model_type == "classifier":
        p = softmax_np(f_out)
        return p * u - p * (p * u).sum(-1, keepdims=True)
    return u


def _sqrt_H_apply_T(model_type, f_out, vec, logvar):
    """ggn.py:16-27  ('sqrt_Hi_apply_T'):  L vec = s*vec - (s.vec) p."""
    if model_type == "regressor":
        return math.sqrt(math.exp(-logvar)) * vec
    p = softmax_np(f_out)
    s = np.sqrt(p)
    return s * vec - (s * vec).sum(-1, keepdims=True) * p


def _sqrt_H_apply(model_type, f_out, vec, logvar):
    """ggn.py:29-39  ('sqrt_Hi_apply'):  L^T vec = s*vec - (p.vec) s."""
    if model_type == "regressor":
        return math.sqrt(math.exp(-logvar)) * vec
    p = softmax_np(f_out)
    s = np.sqrt(p)
    return s * vec - (p * vec).sum(-1, keepdims=True) * s


def compute_ggn_vp(state: OracleState, Z, model_type, full_set_size=None, *, sequential=False,
                   dtype=F64) -> Callable:
    """ggn.py:97-146.  v[D] -> (N/M) * sum_i J_i^T H_i J_i v  (x exp(-logvar) for regressors).

    sequential=True mirrors the reference structure literally (fori_loop over points, per-point jvp,
    redundant forward, per-point vjp: ggn.py:133-144); the default does the same arithmetic batched
    over points (identical up to summation order)."""
    theta, _ = state.flat()
    th = _t(theta, dtype)
    Z = np.asarray(Z)
    M = Z.shape[0]
    N = full_set_size or M
    recal = N / M
    if model_type == "regressor":
        recal *= math.exp(-state.logvar)

    def ggn_vp(v):
        vt = _t(v, dtype)
        if sequential:
            acc = torch.zeros_like(th)
            for i in range(M):
                fzi = lambda p: state.f_theta(p, _t(Z[i:i + 1], dtype))[0].squeeze()
                _, jv = torch.func.jvp(fzi, (th,), (vt,))
                f_val = fzi(th)
                hv = _t(_H_action(model_type, f_val.detach().numpy().astype(np.float64),
                                  jv.detach().numpy().astype(np.float64)), dtype)
                _, vjp_fn = torch.func.vjp(fzi, th)
                acc = acc + vjp_fn(hv)[0]
            return (acc * recal).detach().numpy()
        f = _model_fn(state, Z, dtype)
        fz, jv = torch.func.jvp(f, (th,), (vt,))
        hv = _t(_H_action(model_type, fz.detach().numpy().astype(np.float64),
                          jv.detach().numpy().astype(np.float64)), dtype)
        _, vjp_fn = torch.func.vjp(f, th)
        return (vjp_fn(hv)[0] * recal).detach().numpy()

    return ggn_vp


def compute_W_vps(state: OracleState, Z, model_type, full_set_size=None, blockwise=False):
    """ggn.py:9-93.  Returns (Wfun: [M,K]->[D], WTfun: [D]->[M,K]) ([M] for regressors)."""
    theta, _ = state.flat()
    th = _t(theta)
    Z = np.asarray(Z)
    M = Z.shape[0]
    N = full_set_size or M
    recal = math.sqrt(N / M)
    lv = state.logvar

    def WT_per_point(i, v):
        fzi = lambda p: state.f_theta(p, _t(Z[i:i + 1]))[0].squeeze()
        f_val, jv = torch.func.jvp(fzi, (th,), (_t(v),))
        return recal * _sqrt_H_apply(model_type, f_val.detach().numpy(), jv.detach().numpy(), lv)

    def W_per_point(i, U_i):
        fzi = lambda p: state.f_theta(p, _t(Z[i:i + 1]))[0].squeeze()
        f_val, vjp_fn = torch.func.vjp(fzi, th)
        h = _sqrt_H_apply_T(model_type, f_val.detach().numpy(), np.asarray(U_i, dtype=np.float64), lv)
        return recal * vjp_fn(_t(h).reshape(f_val.shape))[0].detach().numpy()

    if blockwise:
        return W_per_point, WT_per_point

    f = _model_fn(state, Z)
    squeeze = model_type == "regressor"

    def WTfun(v):
        fz, jv = torch.func.jvp(f, (th,), (_t(v),))
        out = recal * _sqrt_H_apply(model_type, fz.detach().numpy(), jv.detach().numpy(), lv)
        return out[:, 0] if squeeze else out  # regressor: (M,) (ggn.py:58,85)

    def Wfun(U):
        U = np.asarray(U, dtype=np.float64)
        if squeeze:
            U = U.reshape(M, 1)
        fz, vjp_fn = torch.func.vjp(f, th)
        h = _sqrt_H_apply_T(model_type, fz.detach().numpy(), U, lv)
        return recal * vjp_fn(_t(h))[0].detach().numpy()

    return Wfun, WTfun


def compute_ggn_dense(state: OracleState, Z, model_type, full_set_size=None):
    """ggn.py:149-193 from explicit Jacobians."""
    J = jacobians(state, Z)  # [M,K,D]
    M = J.shape[0]
    D = J.shape[2]
    G = np.zeros((D, D))
    if model_type == "classifier":
        P = softmax_np(model_outputs(state, Z))
        for i in range(M):
            H = np.diag(P[i]) - np.outer(P[i], P[i])
            G += J[i].T @ H @ J[i]
    else:
        for i in range(M):
            G += J[i].T @ J[i]
        G *= math.exp(-state.logvar)
    N = full_set_size or M
    G *= N / M
    theta, unravel = state.flat()
    return G, theta, unravel


def build_WTW(W, WT, inner_shape, d, *, dtype=np.float64, block=64):
    """ggn.py:198-227: columns = WT(W(one_hot)), symmetrised from the upper triangle (:227)."""
    G = np.zeros((d, d), dtype=dtype)
    for j in range(d):
        e = np.zeros(d, dtype=dtype)
        e[j] = 1.0
        G[:, j] = np.asarray(WT(W(e.reshape(inner_shape)))).reshape(-1)
    return np.triu(G) + np.triu(G, 1).T


def ensure_symmetry(Mx, jitter=1e-8):
    """ggn.py:277"""
    return 0.5 * (Mx + Mx.T) + jitter * np.eye(Mx.shape[0])


# ======================================================================================
# src/lla.py
# ======================================================================================
def compute_curvature_approx(map_state, Z, model_type, alpha, full_set_size=None, **kw):
    """lla.py:11-23"""
    ggn_vp = compute_ggn_vp(map_state, Z, model_type, full_set_size, **kw)

    def curvature_vp(v):
        return ggn_vp(v) + alpha * np.asarray(v, dtype=np.float64)

    return curvature_vp


def compute_curvature_approx_dense(map_state, x, model_type, alpha, full_set_size=None):
    """lla.py:26-34"""
    G, theta, unravel = compute_ggn_dense(map_state, x, model_type, full_set_size)
    return G + alpha * np.eye(G.shape[0]), theta, unravel


def materialize_covariance(f_cov_vp, N, out_dim, mode="diag"):
    """lla.py:160-217"""
    K = N * out_dim
    if mode == "diag":
        diag = np.zeros(K)
        for i in range(K):
            e = np.zeros(K)
            e[i] = 1.0
            diag[i] = np.asarray(f_cov_vp(e)).reshape(K)[i]
        return diag.reshape(N, out_dim)
    if mode == "full":
        cov = np.zeros((K, K))
        for i in range(K):
            e = np.zeros(K)
            e[i] = 1.0
            cov[:, i] = np.asarray(f_cov_vp(e)).reshape(K)
        return cov
    raise ValueError("mode must be 'diag' or 'full'")


# ======================================================================================
# jax.scipy.sparse.linalg.cg  (third-party, absent; restated from its published algorithm)
# ======================================================================================
def cg(A: Callable, b, x0=None, *, tol=1e-5, atol=0.0, maxiter=None, dtype=np.float64):
    """x0=0; stop when r.r <= max(tol^2 b.b, atol^2) or k == maxiter (default 10*n); M = identity.
    Call sites: stochtrace.py:146,192; sample.py:71.
    dtype=np.float32 runs the SAME recurrence in the reference's production precision (main.py:28): every vector, dot product
    and scalar is float32, which is what shows the attainable accuracy of the reference's algorithm on ill-conditioned operators
    (tests/test_gpu_config_parity.py::test_C2_*)."""
    ft = np.dtype(dtype).type
    b = np.asarray(b, dtype=dtype)
    n = b.size
    maxiter = 10 * n if maxiter is None else maxiter
    x = np.zeros_like(b) if x0 is None else np.asarray(x0, dtype=dtype).copy()
    r = (b - np.asarray(A(x), dtype=dtype)) if x0 is not None else b.copy()
    p = r.copy()
    gamma = ft(r.ravel() @ r.ravel())
    atol2 = max(ft(tol) * ft(tol) * ft(b.ravel() @ b.ravel()), ft(atol) * ft(atol))
    k = 0
    while gamma > atol2 and k < maxiter:
        Ap = np.asarray(A(p), dtype=dtype)
        a = ft(gamma / ft(p.ravel() @ Ap.ravel()))
        x = x + a * p
        r = r - a * Ap
        gamma_ = ft(r.ravel() @ r.ravel())
        p = r + ft(gamma_ / gamma) * p
        gamma = gamma_
        k += 1
    return x, k


# ======================================================================================
# matfree (third-party, absent): decomp.tridiag_sym / decomp.bidiag / funm.* / stochtrace.estimator
# ======================================================================================
def tridiag_sym(num_matvecs: int):
    """matfree.decomp.tridiag_sym(k) with full re-orthogonalisation (call sites sample.py:114,
    tests/test_sample.py:338).  Arnoldi/Hessenberg form: for i<k: q_i = v/|v|; v = A q_i; h = Q^T v;
    v -= Q h; v -= Q (Q^T v); H[:,i] = h, H[i+1,i] = |v|.  T = (H + H^T)/2 restricted to its three
    diagonals.  Returns (Q[n,k], T[k,k])."""
    k = num_matvecs

    def decompose(matvec, vec):
        v = np.asarray(vec, dtype=np.float64).copy()
        n = v.size
        if k > n:
            raise ValueError(f"num_matvecs={k} exceeds the operator dimension {n}")
        Q = np.zeros((n, k))
        H = np.zeros((k, k))
        length = np.linalg.norm(v)
        for i in range(k):
            v = v / length
            Q[:, i] = v
            v = np.asarray(matvec(v), dtype=np.float64)
            h = Q.T @ v
            v = v - Q @ h
            v = v - Q @ (Q.T @ v)
            length = np.linalg.norm(v)
            H[:, i] = h
            if i + 1 < k:
                H[i + 1, i] = length
        T = 0.5 * (H + H.T)
        diag = np.diagonal(T).copy()
        off = np.diagonal(T, 1).copy()
        return Q, np.diag(diag) + np.diag(off, 1) + np.diag(off, -1)

    return decompose


def bidiag(num_matvecs: int):
    """matfree.decomp.bidiag(k): Golub-Kahan-Lanczos with full re-orthogonalisation
    (call site train_inducing.py:156).  Returns (U[k,nrows], B[k,k] upper-bidiagonal, V[k,ncols])."""
    k = num_matvecs

    def nrm(x):
        l = np.linalg.norm(x)
        return x / l, l

    def decompose(Av, vA, v0):
        v0 = np.asarray(v0, dtype=np.float64)
        ncols = v0.size
        vk, _ = nrm(v0)
        vk, _ = nrm(vk)
        nrows = np.asarray(Av(vk)).size
        Us = np.zeros((k, nrows))
        Vs = np.zeros((k, ncols))
        alphas = np.zeros(k)
        betas = np.zeros(k)
        beta = 0.0
        for i in range(k):
            Vs[i] = vk
            betas[i] = beta
            uk = np.asarray(Av(vk), dtype=np.float64) - beta * Us[i - 1]
            uk, alpha = nrm(uk)
            uk = uk - Us.T @ (Us @ uk)
            uk, _ = nrm(uk)
            Us[i] = uk
            alphas[i] = alpha
            vk = np.asarray(vA(uk), dtype=np.float64) - alpha * Vs[i]
            vk, beta = nrm(vk)
            vk = vk - Vs.T @ (Vs @ vk)
            vk, _ = nrm(vk)
        B = np.diag(alphas) + np.diag(betas[1:], 1)
        return Us, B, Vs

    return decompose


def dense_funm_sym_eigh(matfun, *, clip_min: Optional[float] = 1.0):
    """matfree_monkeypatch.py:8-22 (clip_min=1.0, line 19); clip_min=None is matfree's own
    (unpatched) dense_funm_sym_eigh used by tests/test_sample.py:9,337."""

    def fun(dense):
        w, V = np.linalg.eigh(dense)
        if clip_min is not None:
            w = np.clip(w, clip_min, None)
        return V @ np.diag(matfun(w)) @ V.T

    return fun


def funm_lanczos_sym(dense_funm, tridiag):
    """matfree.funm.funm_lanczos_sym: f(A)v ~= |v| Q f(T) e1  (call site sample.py:115)."""

    def estimate(matvec, vec):
        vec = np.asarray(vec, dtype=np.float64)
        length = np.linalg.norm(vec)
        Q, T = tridiag(matvec, vec / length)
        fT = dense_funm(T)
        return length * (Q @ fT[:, 0])

    return estimate


def integrand_funm_sym(dense_funm, tridiag):
    """matfree.funm.integrand_funm_sym: v -> |v|^2 e1^T f(T) e1."""

    def quadform(matvec, v0):
        v0 = np.asarray(v0, dtype=np.float64).ravel()
        length = np.linalg.norm(v0)
        _, T = tridiag(matvec, v0 / length)
        return length ** 2 * dense_funm(T)[0, 0]

    return quadform


def integrand_funm_sym_logdet(tridiag, *, clip_min: Optional[float] = 1.0):
    """matfree_monkeypatch.py:25-41 (patched: eigenvalues clipped to >= 1 before log)."""
    return integrand_funm_sym(dense_funm_sym_eigh(np.log, clip_min=clip_min), tridiag)


def dense_funm_product_svd(matfun):
    """matfree.funm.dense_funm_product_svd: B -> V f(S^2) V^T  (no clip; matfree's own)."""

    def fun(B):
        _, S, Vt = np.linalg.svd(B, full_matrices=False)
        return Vt.T @ (matfun(S ** 2)[:, None] * Vt)

    return fun


def integrand_funm_product_logdet(bidiag_alg):
    """matfree.funm.integrand_funm_product_logdet (call site train_inducing.py:157):
    v -> |v|^2 e1^T V log(S^2) V^T e1 with B = U S V^T the GKL bidiagonal of A; A^T via vjp."""
    dense = dense_funm_product_svd(np.log)

    def quadform(Av, vA, v0):
        v0 = np.asarray(v0, dtype=np.float64).ravel()
        length = np.linalg.norm(v0)
        _, B, _ = bidiag_alg(Av, vA, v0 / length)
        return length ** 2 * dense(B)[0, 0]

    return quadform


def estimator(integrand, probes):
    """matfree.stochtrace.estimator with a sampler that ignores its key
    (train_inducing.py:141-142): mean over probe rows."""

    def estimate(*ops):
        return float(np.mean([integrand(*ops, p) for p in probes]))

    return estimate


# ======================================================================================
# src/stochtrace.py   (probe matrices are INPUTS: jax.random is not reproducible without JAX)
# ======================================================================================
def stochastic_trace_estimator_dense(X, eps):
    """stochtrace.py:7-19"""
    return float(np.mean([e @ (X @ e) for e in eps]))


def stochastic_trace_estimator_mvp(Xfun, eps):
    """stochtrace.py:22-34 : mean_b eps_b . X eps_b"""
    return float(np.mean([e @ np.asarray(Xfun(e)) for e in eps]))


def hutchpp_dense(X, eps):
    """stochtrace.py:37-49 ; eps [2*num_samples, n]"""
    ns = eps.shape[0] // 2
    S, G = eps[:ns], eps[ns:]
    Q, _ = np.linalg.qr(X @ S.T)
    P = np.eye(Q.shape[0]) - Q @ Q.T
    return float(np.trace(Q.T @ X @ Q) + np.trace(G @ P @ X @ P @ G.T) / ns)


def hutchpp_mvp(Xfun, eps):
    """stochtrace.py:52-79 ; Xfun takes a MATRIX [n,k] (:64,74)."""
    ns = eps.shape[0] // 2
    S, G = eps[:ns], eps[ns:]
    Q, _ = np.linalg.qr(np.asarray(Xfun(S.T)))
    P = np.eye(Q.shape[0]) - Q @ Q.T
    quad = lambda Mx: Mx.T @ np.asarray(Xfun(Mx))
    return float(np.trace(quad(Q)) + np.trace(quad(P @ G.T)) / ns)


def hutchpp(Xfun, eps):
    """stochtrace.py:82-111 ; Xfun takes a VECTOR; note the 1/num_samples uses the FULL probe count (:84,109)."""
    num_samples = eps.shape[0]
    S, G = eps[:num_samples // 2], eps[num_samples // 2:]
    Y = np.stack([np.asarray(Xfun(s)) for s in S], axis=1)
    Q, _ = np.linalg.qr(Y)
    P = np.eye(Q.shape[0]) - Q @ Q.T

    def quad(Mx):
        Yx = np.stack([np.asarray(Xfun(Mx[:, j])) for j in range(Mx.shape[1])], axis=1)
        return Mx.T @ Yx

    return float(np.trace(quad(Q)) + np.trace(quad(P @ G.T)) / num_samples)


def apply_X(Xfun, Mx):
    """stochtrace.py:113-114 : rows are probes -> columns of the result"""
    return np.stack([np.asarray(Xfun(r)) for r in Mx], axis=1)


def hutchpp_v2(Xfun, eps, *, s1, s2):
    """stochtrace.py:118-135"""
    S, G = eps[:s1], eps[s1:]
    Y = apply_X(Xfun, S)
    Q, _ = np.linalg.qr(Y)
    XQ = apply_X(Xfun, Q.T)
    low_rank = np.trace(XQ.T @ Q)
    G_perp = G - (G @ Q) @ Q.T
    XGp = apply_X(Xfun, G_perp)
    resid = np.trace(G_perp @ XGp) / s2
    return float(low_rank + resid)


def hutchpp_inv_mvp(Xfun, eps):
    """stochtrace.py:138-148 ; Xfun takes a vector; CG is applied column-wise."""
    def Xinv(Mx):
        Mx = np.asarray(Mx)
        if Mx.ndim == 1:
            return cg(Xfun, Mx)[0]
        return np.stack([cg(Xfun, Mx[:, j])[0] for j in range(Mx.shape[1])], axis=1)
    return hutchpp_mvp(Xinv, eps)


def na_hutchpp_dense(X, eps):
    """stochtrace.py:151-163 ; eps [4*num_samples, n]"""
    ns = eps.shape[0] // 4
    c3 = 0.25
    S, R, G = eps[:ns], eps[ns:3 * ns], eps[3 * ns:]
    W = X @ S.T
    Zm = X @ R.T
    pin = np.linalg.pinv(S @ Zm)
    return float(np.trace(pin @ (W.T @ Zm))
                 + (np.trace(G @ X @ G.T) - np.trace(G @ Zm @ pin @ W.T @ G.T)) / (c3 * 4 * ns))


def na_hutchpp_mvp(Xfun, eps):
    """stochtrace.py:166-180 ; Xfun takes a matrix"""
    ns = eps.shape[0] // 4
    c3 = 0.25
    S, R, G = eps[:ns], eps[ns:3 * ns], eps[3 * ns:]
    W = np.asarray(Xfun(S.T))
    Zm = np.asarray(Xfun(R.T))
    pin = np.linalg.pinv(S @ Zm)
    return float(np.trace(pin @ (W.T @ Zm))
                 + (np.trace(G @ np.asarray(Xfun(G.T))) - np.trace(G @ Zm @ pin @ W.T @ G.T)) / (c3 * 4 * ns))


def na_hutchpp_inv_mvp(Xfun, eps):
    """stochtrace.py:183-194"""
    def Xinv(Mx):
        Mx = np.asarray(Mx, dtype=np.float64)
        if Mx.ndim == 1:
            return cg(Xfun, Mx)[0]
        return np.stack([cg(Xfun, Mx[:, j])[0] for j in range(Mx.shape[1])], axis=1)
    return na_hutchpp_mvp(Xinv, eps)


# ======================================================================================
# train_inducing.py:148-171  — the production SLQ logdet (GKL form) and the Lanczos form
# ======================================================================================
def slq_logdet_gkl(state, Z, model_type, alpha, probes, num_matvecs):
    """logdet(alpha I + Wz Wz^T) via GKL on A = [sqrt(alpha) I; Wz^T]  (train_inducing.py:114-116,
    156-171; NB full_set_size=None inside compute_W_vps, i.e. no beta)."""
    Wz, WzT = compute_W_vps(state, Z, model_type, full_set_size=None)
    sa = math.sqrt(alpha)
    D = state.flat()[0].size
    inner = np.asarray(WzT(np.zeros(D))).shape

    def Av(v):
        return np.concatenate([sa * v, np.asarray(WzT(v)).ravel()])

    def vA(u):
        return sa * u[:D] + Wz(u[D:].reshape(inner))

    integrand = integrand_funm_product_logdet(bidiag(num_matvecs))
    return estimator(integrand, probes)(Av, vA)


def alternative_objective_scalable(Z, X, state, alpha, model_type, probes, full_set_size=None, st_samples=None,
                                   slq_samples=2, slq_num_matvecs=None):
    """train_inducing.py:87-173 (forward value): tr(S_X S_Z^{-1}) by hutchpp_v2 on the Woodbury inverse + the GKL logdet.
    `probes` [st_samples, D] replaces the matfree Rademacher sampler (:138-142; the same probes feed both terms)."""
    import scipy.linalg as sla
    probes = np.asarray(probes, dtype=np.float64)
    st_samples = probes.shape[0] if st_samples is None else st_samples
    N = full_set_size
    M = np.asarray(Z).shape[0]
    beta = N / M
    S_vp = compute_curvature_approx(state, X, model_type, alpha, full_set_size=N)                # :108-110
    Wz, WzT = compute_W_vps(state, Z, model_type, full_set_size=None)                            # :114-116
    D = state.flat()[0].size
    inner = np.asarray(WzT(np.zeros(D))).shape
    d_z = int(np.prod(inner))
    WzTWz = build_WTW(Wz, WzT, inner, d_z, block=1)                                              # :126
    Kmat = np.eye(d_z) / beta + WzTWz / alpha

    def Sz_inv(v):                                                                               # :127-132
        u = np.asarray(WzT(v)).reshape(d_z)
        x = sla.solve(Kmat, u)
        return v / alpha - Wz(x.reshape(inner)) / alpha ** 2

    trace_term = hutchpp_v2(lambda v: S_vp(Sz_inv(v)), probes[:st_samples], s1=st_samples - 16, s2=16)   # :144-145
    k = slq_num_matvecs if slq_num_matvecs is not None else int(M * 0.8)                         # :148
    logdet_term = slq_logdet_gkl(state, Z, model_type, alpha, probes[:slq_samples], k)           # :156-171
    return logdet_term + trace_term


def slq_logdet_lanczos(matvec, probes, num_matvecs, *, clip_min=1.0):
    """train_inducing.py:152-153 (commented 'old tridiagonal formulation') / tests/test_variational.py:126-150."""
    integrand = integrand_funm_sym_logdet(tridiag_sym(num_matvecs), clip_min=clip_min)
    return estimator(integrand, probes)(matvec)


# ======================================================================================
# src/sample.py
# ======================================================================================
def inv_matsqrt_dense(state, Z, alpha, model_type, full_set_size=None):
    """sample.py:16-52 (debug twin)."""
    theta, _ = state.flat()
    D = theta.size
    M = np.asarray(Z).shape[0]
    N = full_set_size or M
    beta = N / M
    Wfun, WTfun = compute_W_vps(state, Z, model_type, full_set_size=None)
    I_D = np.eye(D)
    W = np.stack([np.asarray(WTfun(I_D[j])).ravel() for j in range(D)], axis=0)  # (D, d)
    WT = W.T
    comp = WT @ W
    inv_comp = np.linalg.solve(comp, np.eye(comp.shape[0]))
    nullproj = I_D - W @ inv_comp @ WT
    term1 = nullproj / math.sqrt(alpha)
    w, V = np.linalg.eigh(alpha * np.eye(comp.shape[0]) + beta * comp)
    inv_sqrt = (V * (1.0 / np.sqrt(np.clip(w, 0, np.inf)))) @ V.T
    return term1 + W @ inv_comp @ inv_sqrt @ WT


def inv_matsqrt_vp(state, Z, D, alpha, model_type, full_set_size=None):
    """sample.py:55-145 with key=None (direct projection; sample.py:150 forces it)."""
    Wfun, WTfun = compute_W_vps(state, Z, model_type, full_set_size=None)
    dummy = np.asarray(WTfun(np.zeros(D)))
    inner_shape, d = dummy.shape, dummy.size
    WTW = build_WTW(Wfun, WTfun, inner_shape, d, dtype=np.float64, block=2)
    import scipy.linalg as sla

    def nullproj_vp(v):
        u = np.asarray(WTfun(v)).ravel()
        x = sla.solve(WTW, u)
        return v - Wfun(x.reshape(inner_shape))

    M = np.asarray(Z).shape[0]
    N = full_set_size or M
    beta = N / M
    invsqrt_fun = dense_funm_sym_eigh(lambda x: 1.0 / np.sqrt(x), clip_min=1.0)  # sample.py:113
    invmatsqrt = funm_lanczos_sym(invsqrt_fun, tridiag_sym(2 * M))              # sample.py:114-115

    def invmatsqrt_term(V):
        Vflat = np.asarray(V).ravel()
        return invmatsqrt(lambda u: alpha * u + beta * (WTW @ u), Vflat).reshape(inner_shape)

    def outer_fun(v):
        u = invmatsqrt_term(WTfun(v)).ravel()
        x = sla.solve(WTW, u)
        return Wfun(x.reshape(inner_shape))

    def vp(v):
        v = np.asarray(v, dtype=np.float64)
        return outer_fun(v) + nullproj_vp(v) / math.sqrt(alpha)

    return vp


def sample(state, Z, D, alpha, Eps, model_type, full_set_size=None):
    """sample.py:148-156 with the N(0,I) draws Eps[S,D] supplied by the caller."""
    f = inv_matsqrt_vp(state, Z, D, alpha, model_type, full_set_size=full_set_size)
    return np.stack([f(e) for e in np.asarray(Eps, dtype=np.float64)])


def predict_lla_scalable(map_state, Xnew, Z, model_type, alpha, Eps, full_set_size=None):
    """lla.py:133-156: f(theta*,X) + J_X w_s for w_s = A^{-1/2} eps_s."""
    theta, _ = map_state.flat()
    D = theta.size
    w_samples = sample(map_state, Z, D, alpha, Eps, model_type, full_set_size=full_set_size)
    f = _model_fn(map_state, Xnew)
    th = _t(theta)
    fmu = f(th).detach().numpy()
    dys = np.stack([torch.func.jvp(f, (th,), (_t(w),))[1].detach().numpy() for w in w_samples])
    return fmu[None] + dys


# ======================================================================================
# gradients with respect to the inducing points Z
# ======================================================================================
# The reference obtains them with jax.value_and_grad THROUGH the closures of src/ggn.py (train_inducing.py:195-232).
# Restated literally: the operators below are written in differentiable torch (float64) exactly as ggn.py composes
# them (jvp -> output-space factor -> vjp) and torch.func.grad differentiates the scalar <cotangent, operator(vector)>
# with respect to Z.  No hand-derived adjoint appears here, so this checks the CUDA path's hand-written reverse pass.
def _softmax_t(f):
    return torch.softmax(f, dim=-1)


def _ggn_vp_t(state, Zt, th, model_type, recal, v):
    """ggn.py:133-144 batched over points, differentiable in Zt."""
    f = lambda p: state.f_theta(p, Zt)
    fz, jv = torch.func.jvp(f, (th,), (v,))
    if model_type == "classifier":
        p = _softmax_t(fz)
        hv = p * jv - p * (p * jv).sum(-1, keepdim=True)      # ggn.py:125-131
    else:
        hv = jv
    _, vjp_fn = torch.func.vjp(f, th)
    return recal * vjp_fn(hv)[0]


def _wt_t(state, Zt, th, model_type, recal, lv, v):
    """ggn.py:54-62,84-85 + sqrt_Hi_apply (ggn.py:29-39)."""
    f = lambda p: state.f_theta(p, Zt)
    fz, jv = torch.func.jvp(f, (th,), (v,))
    if model_type == "regressor":
        return recal * math.sqrt(math.exp(-lv)) * jv
    p = _softmax_t(fz)
    s = torch.sqrt(p)
    return recal * (s * jv - (p * jv).sum(-1, keepdim=True) * s)


def _w_t(state, Zt, th, model_type, recal, lv, U):
    """ggn.py:64-76,87-91 + sqrt_Hi_apply_T (ggn.py:16-27)."""
    f = lambda p: state.f_theta(p, Zt)
    fz, vjp_fn = torch.func.vjp(f, th)
    if model_type == "regressor":
        h = math.sqrt(math.exp(-lv)) * U.reshape(fz.shape)
    else:
        p = _softmax_t(fz)
        s = torch.sqrt(p)
        h = s * U - (s * U).sum(-1, keepdim=True) * p
    return recal * vjp_fn(h)[0]


def ggn_vp_zgrad(state: OracleState, Z, model_type, Ubar, V, full_set_size=None, per_probe=False) -> np.ndarray:
    """d/dZ sum_b <Ubar[b], ggn_vp_Z(V[b])>  -> [M, ...] (or [B, M, ...])."""
    theta, _ = state.flat()
    th = _t(theta)
    Zt = _t(Z)
    M = Zt.shape[0]
    recal = (full_set_size or M) / M
    if model_type == "regressor":
        recal *= math.exp(-state.logvar)
    Ub, Vb = _t(Ubar).reshape(-1, th.numel()), _t(V).reshape(-1, th.numel())
    grads = []
    for u, v in zip(Ub, Vb):
        grads.append(torch.func.grad(lambda Zv: (u * _ggn_vp_t(state, Zv, th, model_type, recal, v)).sum())(Zt).numpy())
    g = np.stack(grads)
    return g if per_probe else g.sum(0)


def W_vps_zgrad(state: OracleState, Z, model_type, full_set_size=None):
    """Returns (W_zgrad(ubar [B,D], U [B,M,K]), WT_zgrad(Ybar [B,M,K], v [B,D])): d/dZ of <ubar, Wfun(U)> and <Ybar, WTfun(v)>."""
    theta, _ = state.flat()
    th = _t(theta)
    Zt = _t(Z)
    M = Zt.shape[0]
    recal = math.sqrt((full_set_size or M) / M)
    lv = state.logvar
    D = th.numel()

    def W_zgrad(ubar, U):
        ub = _t(ubar).reshape(-1, D)
        Ub = _t(U).reshape(ub.shape[0], M, -1)
        return sum(torch.func.grad(lambda Zv: (u * _w_t(state, Zv, th, model_type, recal, lv, Uu)).sum())(Zt).numpy()
                   for u, Uu in zip(ub, Ub))

    def WT_zgrad(Ybar, v):
        vb = _t(v).reshape(-1, D)
        Yb = _t(Ybar).reshape(vb.shape[0], M, -1)
        return sum(torch.func.grad(lambda Zv: (Y * _wt_t(state, Zv, th, model_type, recal, lv, vv)).sum())(Zt).numpy()
                   for Y, vv in zip(Yb, vb))

    return W_zgrad, WT_zgrad


def jvp_zgrad(state: OracleState, Z, Cbar, V) -> np.ndarray:
    """d/dZ sum_b <Cbar[b], J_Z V[b]>   (the plain batched JVP of lla.py:153)."""
    theta, _ = state.flat()
    th = _t(theta)
    Zt = _t(Z)
    vb = _t(V).reshape(-1, th.numel())
    Cb = _t(Cbar).reshape(vb.shape[0], Zt.shape[0], -1)
    return sum(torch.func.grad(lambda Zv: (Cm * torch.func.jvp(lambda p: state.f_theta(p, Zv), (th,), (vv,))[1]).sum())(Zt).numpy()
               for Cm, vv in zip(Cb, vb))


# ======================================================================================
# deterministic objectives and their gradients with respect to Z (train_inducing.py:26-84,175-192; train_alpha.py:13-44)
# ======================================================================================
def _jac_t(state, Zt, th):
    """J[i] = d f(z_i; theta) / d theta, [M, K, D], differentiable in Zt."""
    return torch.func.jacrev(lambda p: state.f_theta(p, Zt))(th)


def _ggn_dense_t(state, Zt, th, model_type, full_set_size):
    """ggn.py:149-193 (compute_ggn_dense) in differentiable torch."""
    M = Zt.shape[0]
    recal = (full_set_size or M) / M
    J = _jac_t(state, Zt, th)
    if model_type == "classifier":
        p = _softmax_t(state.f_theta(th, Zt))
        H = torch.diag_embed(p) - p[:, :, None] * p[:, None, :]
        return recal * torch.einsum("ikd,ikl,ile->de", J, H, J)
    return recal * math.exp(-state.logvar) * torch.einsum("ikd,ike->de", J, J)


def _w_dense_t(state, Zt, th, model_type):
    """The matrix W = [J_1^T L_1 ... J_M^T L_M] in R^{D x MK} (ggn.py:79-93, full_set_size=None), differentiable in Zt."""
    J = _jac_t(state, Zt, th)
    M, K, D = J.shape
    if model_type == "classifier":
        p = _softmax_t(state.f_theta(th, Zt))
        r = torch.sqrt(p)
        L = torch.diag_embed(r) - p[:, :, None] * r[:, None, :]          # L u = r*u - (r.u) p   (ggn.py:23-27)
        return torch.einsum("ikd,ikl->dil", J, L).reshape(D, M * K)
    return math.sqrt(math.exp(-state.logvar)) * J.permute(2, 0, 1).reshape(D, M * K)


def _objective_dense_t(state, Zt, Xt, th, alpha, model_type, full_set_size):
    """train_inducing.py:175-192 (alternative_objective_dense)."""
    D = th.numel()
    eye = torch.eye(D, dtype=F64)
    S = _ggn_dense_t(state, Xt, th, model_type, full_set_size) + alpha * eye              # lla.py compute_curvature_approx_dense
    S_z = _ggn_dense_t(state, Zt, th, model_type, full_set_size) + alpha * eye
    S_z_inv = torch.linalg.inv(S_z)
    trace_term = torch.trace(S @ S_z_inv)
    return -torch.linalg.slogdet(S_z_inv)[1] + trace_term


def _objective_exact_t(state, Zt, Xt, th, alpha, model_type, full_set_size):
    """train_inducing.py:26-84 (alternative_objective_scalable_exact)."""
    N = full_set_size
    M, Kx = Zt.shape[0], Xt.shape[0]
    beta, gamma = N / M, N / Kx
    D = th.numel()
    Wz = _w_dense_t(state, Zt, th, model_type)
    Wx = _w_dense_t(state, Xt, th, model_type)
    WzTWz = Wz.T @ Wz                                                                     # :60
    d_z = WzTWz.shape[0]
    eye = torch.eye(d_z, dtype=F64)
    logdet_term = torch.linalg.slogdet(eye + beta / alpha * WzTWz)[1] + D * math.log(alpha)   # :62-63
    WTWz = Wx.T @ Wz                                                                      # :67
    Mm = eye / beta + WzTWz / alpha                                                       # :69
    L = torch.linalg.cholesky(Mm)
    S1 = torch.cholesky_solve(WzTWz, L)                                                   # :71
    S2 = torch.cholesky_solve(WTWz.T, L)                                                  # :72
    trace1 = torch.trace(S1)
    trace2 = (WTWz * S2.T).sum()                                                          # :75
    return logdet_term - trace1 / alpha - gamma / alpha ** 2 * trace2                     # :76-78


def _value_and_zgrad(fn, state, Z, X, alpha, model_type, full_set_size):
    theta, _ = state.flat()
    th = _t(theta)
    Zt, Xt = _t(Z), _t(X)
    g, v = torch.func.grad_and_value(lambda Zv: fn(state, Zv, Xt, th, alpha, model_type, full_set_size))(Zt)
    return float(v), g.numpy()


def variational_grad_dense(Z, X, state, alpha, model_type, full_set_size=None):
    """jax.value_and_grad(alternative_objective_dense) (train_inducing.py:194) -> (loss, dLoss/dZ)."""
    return _value_and_zgrad(_objective_dense_t, state, Z, X, alpha, model_type, full_set_size)


def variational_grad_scalable_exact(Z, X, state, alpha, model_type, full_set_size=None):
    """value and Z-gradient of alternative_objective_scalable_exact (train_inducing.py:26-84)."""
    return _value_and_zgrad(_objective_exact_t, state, Z, X, alpha, model_type, full_set_size)


def log_marginal_likelihood(alpha, X, state, model_type, full_set_size=None):
    """train_alpha.py:13-44 -> (log p(D|alpha) up to constants, d/d log(alpha) of it) (update_alpha differentiates in log alpha)."""
    theta, _ = state.flat()
    th = _t(theta)
    Xt = _t(X)
    N = full_set_size or Xt.shape[0]
    rescale = N / Xt.shape[0]
    D = th.numel()
    W = _w_dense_t(state, Xt, th, model_type).detach()
    WTW = W.T @ W
    d = WTW.shape[0]

    def lml(log_alpha):
        a = torch.exp(log_alpha)
        logdet_term = torch.linalg.slogdet(torch.eye(d, dtype=F64) + rescale / a * WTW)[1] + D * torch.log(a)
        log_prior = -0.5 * a * (th @ th) + 0.5 * D * torch.log(a)
        return log_prior - 0.5 * logdet_term

    la = torch.tensor(math.log(alpha), dtype=F64)
    g, v = torch.func.grad_and_value(lml)(la)
    return float(v), float(g)


# ======================================================================================
# evaluation consumer (scale_experiments/evaluate.py:40-152)
# ======================================================================================
def mc_softmax_predictive(logit_samples, y):
    """evaluate.py:119-142: (log( 1/S sum_s softmax(l_s)[y] ) per example, mean_s softmax(l_s))."""
    ls = np.asarray(logit_samples, dtype=np.float64)
    y = np.asarray(y).reshape(-1).astype(np.int64)
    sh = ls - ls.max(axis=-1, keepdims=True)
    logp = sh - np.log(np.exp(sh).sum(axis=-1, keepdims=True))                 # log_softmax, (S,B,C)
    lpt = np.take_along_axis(logp, y[None, :, None], axis=-1)[..., 0]          # (S,B)
    mx = lpt.max(axis=0)
    log_avg = mx + np.log(np.exp(lpt - mx).sum(axis=0)) - math.log(ls.shape[0])
    return log_avg, np.exp(logp).mean(axis=0)


def batch_nll(state, x, y, Z, *, alpha, full_set_size, model_type, Eps):
    """evaluate.py:98-152 with the posterior noise Eps [S, D] as an input -> (nll, acc, mean probs)."""
    logits = predict_lla_scalable(state, x, Z, model_type, alpha, Eps, full_set_size=full_set_size)
    log_avg, mean = mc_softmax_predictive(logits, y)
    acc = float((mean.argmax(-1) == np.asarray(y).reshape(-1)).mean())
    return float(-log_avg.mean()), acc, mean


def brier_score(probs, labels):
    """evaluate.py:40-43."""
    probs = np.asarray(probs, dtype=np.float64)
    onehot = np.zeros_like(probs)
    onehot[np.arange(len(labels)), np.asarray(labels).astype(int)] = 1.0
    return float(((probs - onehot) ** 2).sum(axis=1).mean())


def ece(probs, labels, n_bins=15):
    """evaluate.py:45-63 (half-open bins [lo, hi) on linspace edges)."""
    probs = np.asarray(probs)
    labels = np.asarray(labels).reshape(-1)
    conf, pred = probs.max(1), probs.argmax(1)
    hit = pred == labels
    edges = np.linspace(0.0, 1.0, n_bins + 1)
    out = 0.0
    for b in range(n_bins):
        m = (conf >= edges[b]) & (conf < edges[b + 1])
        if m.any():
            out += abs(conf[m].mean() - hit[m].mean()) * m.mean()
    return float(out)
