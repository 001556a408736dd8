"""GPU parity: the CUDA path (through the C ABI) against the float64 CPU oracle on identical weights, points
and probe matrices.  Tolerances are BASELINE.json's: GGN-vector products <= 1e-5 relative (fp32),
trace / logdet / CG <= 1e-4 relative."""
import math

import numpy as np
import pytest
import torch

from helpers import make_pair, rel_err
from oracle import lip_oracle as O

pytestmark = pytest.mark.gpu

TOL_GGN = 1e-5
TOL_EST = 1e-4

CONFIGS = {
    # name: (kind, hidden, n_out, in_dim, M, N, logvar)
    "C1_toy_sine": ("regressor", [8, 8, 8, 8], 1, 1, 40, 240, 0.0),
    "C2_xor": ("classifier", [16, 16], 2, 2, 32, 800, 0.0),
    "C3a_subset89": ("classifier", [32, 32, 32], 2, 2, 96, 10000, 0.0),
    "mlp_ragged": ("large", [37, 129, 20], 7, 45, 133, 1000, 0.0),
    "reg_logvar": ("regressor", [9], 1, 3, 17, 170, 0.4),
    "linear": ("regressor", [], 1, 1, 4, 4, 0.07),
}


def _setup(name, seed=0):
    kind, hidden, n_out, in_dim, M, N, logvar = CONFIGS[name]
    ost, lst = make_pair(kind, hidden=hidden, n_out=n_out, in_dim=in_dim, seed=100 + seed, logvar=logvar)
    rng = np.random.default_rng(1000 + seed)
    Z = rng.standard_normal((M, in_dim)).astype(np.float32)
    mt = "regressor" if kind == "regressor" else "classifier"
    return ost, lst, Z, mt, N


def cu(x):
    return torch.as_tensor(np.asarray(x), dtype=torch.float32, device="cuda")


@pytest.mark.parametrize("name", list(CONFIGS))
def test_ggn_vp_matches_oracle(name):
    from lip_b200 import ggn, lla
    ost, lst, Z, mt, N = _setup(name)
    D = ost.flat()[0].size
    rng = np.random.default_rng(5)
    V = rng.choice([-1.0, 1.0], size=(6, D)).astype(np.float32)
    V[3:] = rng.standard_normal((3, D)).astype(np.float32)
    ref_vp = O.compute_ggn_vp(ost, Z, mt, full_set_size=N)
    ref = np.stack([ref_vp(v) for v in V.astype(np.float64)])
    vp = ggn.compute_ggn_vp(lst, cu(Z), mt, full_set_size=N)
    got = vp(cu(V)).cpu().numpy()
    assert got.shape == (6, D)
    assert rel_err(got, ref) < TOL_GGN
    single = vp(cu(V[0])).cpu().numpy()          # the reference's un-batched signature
    assert single.shape == (D,) and rel_err(single, ref[0]) < TOL_GGN
    alpha = 0.37
    cvp = lla.compute_curvature_approx(lst, cu(Z), mt, alpha, full_set_size=N)
    assert rel_err(cvp(cu(V)).cpu().numpy(), ref + alpha * V) < TOL_GGN


def test_G1_golden_linear_model():
    """tests/test_ggn.py:87-102 / fixtures.py:24: GGN = e^{-logvar} [[n, sum x],[sum x, sum x^2]] in (bias, kernel) order."""
    from lip_b200 import ggn
    ost, lst, _, mt, _ = _setup("linear")
    X = np.array([[-1.0], [0.0], [1.1], [3.5]], dtype=np.float32)
    G, flat, _ = ggn.compute_ggn_dense(lst, cu(X), "regressor")
    expect = math.exp(-0.07) * np.array([[4.0, 3.6], [3.6, 14.46]])
    np.testing.assert_allclose(G.cpu().numpy(), expect, rtol=2e-6, atol=1e-6)
    Wfun, WTfun = ggn.compute_W_vps(lst, cu(X), "regressor")
    WT_out = WTfun(torch.eye(2, device="cuda"))
    assert WT_out.shape == (2, 4)                  # regressor: (M,) per vector
    comp = Wfun(WT_out)
    np.testing.assert_allclose(comp.cpu().numpy(), expect, rtol=2e-6, atol=1e-6)


@pytest.mark.parametrize("name", ["C1_toy_sine", "C2_xor", "mlp_ragged", "reg_logvar"])
def test_W_WT_match_oracle(name):
    from lip_b200 import ggn
    ost, lst, Z, mt, N = _setup(name)
    D = ost.flat()[0].size
    rng = np.random.default_rng(6)
    V = rng.standard_normal((4, D)).astype(np.float32)
    Wo, WTo = O.compute_W_vps(ost, Z, mt, full_set_size=N)
    Wg, WTg = ggn.compute_W_vps(lst, cu(Z), mt, full_set_size=N)
    ref_wt = np.stack([WTo(v) for v in V.astype(np.float64)])
    got_wt = WTg(cu(V)).cpu().numpy()
    assert got_wt.shape == ref_wt.shape
    assert rel_err(got_wt, ref_wt) < TOL_GGN
    U = rng.standard_normal(ref_wt.shape).astype(np.float32)
    ref_w = np.stack([Wo(u) for u in U.astype(np.float64)])
    got_w = Wg(cu(U)).cpu().numpy()
    assert rel_err(got_w, ref_w) < TOL_GGN
    assert rel_err(Wg(cu(U[0])).cpu().numpy(), ref_w[0]) < TOL_GGN
    # identity W(W^T v) == GGN v  (tests/test_sample.py:19-49)
    vp = ggn.compute_ggn_vp(lst, cu(Z), mt, full_set_size=N)
    assert rel_err(Wg(WTg(cu(V))).cpu().numpy(), vp(cu(V)).cpu().numpy()) < 2e-5
    # blockwise closures (ggn.py:79-82)
    Wb, WTb = ggn.compute_W_vps(lst, cu(Z), mt, full_set_size=N, blockwise=True)
    Wob, WTob = O.compute_W_vps(ost, Z, mt, full_set_size=N, blockwise=True)
    assert rel_err(WTb(3, cu(V[0])).cpu().numpy(), WTob(3, V[0].astype(np.float64))) < TOL_GGN


def test_gram_matches_oracle():
    from lip_b200 import ggn
    ost, lst, Z, mt, N = _setup("C2_xor")
    Wo, WTo = O.compute_W_vps(ost, Z, mt)
    d = Z.shape[0] * 2
    ref = O.build_WTW(Wo, WTo, (Z.shape[0], 2), d)
    Wg, WTg = ggn.compute_W_vps(lst, cu(Z), mt)
    got = ggn.build_WTW(Wg, WTg, (Z.shape[0], 2), d, dtype=torch.float32, block=2).cpu().numpy()
    assert rel_err(got, ref) < TOL_GGN
    np.testing.assert_array_equal(got, got.T)
    # generic (closure-pushing) path gives the same matrix
    plainW = lambda U: Wg(U)
    got2 = ggn.build_WTW(plainW, WTg, (Z.shape[0], 2), d, dtype=torch.float32, block=16).cpu().numpy()
    assert rel_err(got2, ref) < TOL_GGN


def test_forward_outputs_match_oracle():
    from lip_b200 import ggn
    for name in ("C1_toy_sine", "mlp_ragged"):
        ost, lst, Z, mt, N = _setup(name)
        bm = ggn._bind(lst, cu(Z), mt)
        assert rel_err(bm.outputs().cpu().numpy(), O.model_outputs(ost, Z)) < 2e-6


# --------------------------------------------------------------------------------- vector stage
def test_cg_matches_oracle():
    from lip_b200 import matfree
    rng = np.random.default_rng(3)
    n = 300
    G = rng.standard_normal((n, n))
    A = (G @ G.T / n + np.eye(n)).astype(np.float32)
    Bv = rng.standard_normal((5, n)).astype(np.float32)
    At = cu(A)
    mv = matfree.batched(lambda X: X @ At)
    x, iters = matfree.cg(mv, cu(Bv))
    for b in range(5):
        xo, ko = O.cg(lambda v: A.astype(np.float64) @ v, Bv[b].astype(np.float64))
        assert rel_err(x[b].cpu().numpy(), xo) < TOL_EST
        assert abs(int(iters[b]) - ko) <= 1
    x1, _ = matfree.cg(lambda v: At @ v, cu(Bv[0]))        # un-batched closure, single vector
    assert rel_err(x1.cpu().numpy(), O.cg(lambda v: A.astype(np.float64) @ v, Bv[0].astype(np.float64))[0]) < TOL_EST


def test_tridiag_funm_vs_eigh():
    from lip_b200 import matfree
    rng = np.random.default_rng(8)
    for k in (1, 2, 5, 40, 300):
        d = (rng.random((3, k)) * 3 + 0.5).astype(np.float32)
        e = (rng.standard_normal((3, max(k - 1, 0))) * 0.7).astype(np.float32)
        for fn, f, clip in (("log", np.log, None), ("invsqrt", lambda x: 1 / np.sqrt(x), 1.0), ("inv", lambda x: 1 / x, None)):
            dd = d.copy() + 3.0   # diagonally dominant -> SPD
            Ts = [np.diag(dd[b].astype(np.float64)) + np.diag(e[b].astype(np.float64), 1) + np.diag(e[b].astype(np.float64), -1)
                  for b in range(3)]
            df = matfree.DenseFunm(fn, clip)
            q = df.quad_e1(cu(dd), cu(e)).cpu().numpy()
            fe = df.apply_e1(cu(dd), cu(e)).cpu().numpy()
            for b in range(3):
                w, V = np.linalg.eigh(Ts[b])
                if clip is not None:
                    w = np.clip(w, clip, None)
                fT = V @ np.diag(f(w)) @ V.T
                assert abs(q[b] - fT[0, 0]) <= 2e-6 * max(1.0, abs(fT[0, 0]))
                assert rel_err(fe[b], fT[:, 0]) < 5e-6


def test_G2_lanczos_invsqrt_golden_and_oracle():
    """tests/test_sample.py:334-355 on the GPU path, and against the oracle's Lanczos."""
    from lip_b200 import matfree
    D = 100
    diag = np.arange(1, D + 1, dtype=np.float64) / D
    dg = cu(diag)
    invsqrt = matfree.dense_funm_sym_eigh(lambda x: 1.0 / torch.sqrt(x))
    f = matfree.funm_lanczos_sym(invsqrt, matfree.decomp.tridiag_sym(20))
    res = f(lambda v: dg * v, torch.ones(D, device="cuda")).cpu().numpy()
    np.testing.assert_allclose(res, 1.0 / np.sqrt(diag), rtol=1e-1)
    fo = O.funm_lanczos_sym(O.dense_funm_sym_eigh(lambda x: 1.0 / np.sqrt(x), clip_min=None), O.tridiag_sym(20))
    assert rel_err(res, fo(lambda v: diag * v, np.ones(D))) < TOL_EST


def test_lanczos_slq_and_gkl_logdet_match_oracle():
    from lip_b200 import ggn, lla, matfree, matfree_monkeypatch
    ost, lst, Z, mt, N = _setup("C2_xor")
    D = ost.flat()[0].size
    alpha = 1.7
    rng = np.random.default_rng(12)
    probes = rng.choice([-1.0, 1.0], size=(3, D)).astype(np.float32)
    # Lanczos form with the patched (clipped) integrand, operator = curvature_vp
    cvp_o = O.compute_curvature_approx(ost, Z, mt, alpha, full_set_size=N)
    ref = O.slq_logdet_lanczos(cvp_o, probes.astype(np.float64), 25, clip_min=1.0)
    cvp = lla.compute_curvature_approx(lst, cu(Z), mt, alpha, full_set_size=N)
    integrand = matfree_monkeypatch.integrand_funm_sym_logdet(matfree.decomp.tridiag_sym(25))
    est = matfree.stochtrace.estimator(integrand, sampler=lambda _: cu(probes))
    got = float(est(cvp, 0))
    assert abs(got - ref) <= TOL_EST * abs(ref)
    # GKL product form (train_inducing.py:148-171)
    ref2 = O.slq_logdet_gkl(ost, Z, mt, alpha, probes.astype(np.float64), 25)
    Wz, WzT = ggn.compute_W_vps(lst, cu(Z), mt, full_set_size=None)
    sa = math.sqrt(alpha)
    d = Z.shape[0] * 2

    @matfree.batched
    def bidiag_target(v):
        v = v.reshape(-1, D)
        return torch.cat([sa * v, WzT(v).reshape(v.shape[0], d)], dim=1)

    @matfree.batched
    def bidiag_target_T(u):
        u = u.reshape(-1, D + d)
        return sa * u[:, :D] + Wz(u[:, D:].reshape(-1, Z.shape[0], 2))

    problem = matfree.funm.integrand_funm_product_logdet(matfree.decomp.bidiag(25))
    est2 = matfree.stochtrace.estimator(lambda Av, s: problem(Av, s, bidiag_target_T), sampler=lambda _: cu(probes))
    est2_fn = matfree.stochtrace.estimator(matfree.batched(lambda Av, s: problem(Av, s, bidiag_target_T)),
                                           sampler=lambda _: cu(probes))
    got2 = float(est2_fn(bidiag_target, 0))
    assert abs(got2 - ref2) <= TOL_EST * abs(ref2)
    assert abs(float(est2(bidiag_target, 0)) - ref2) <= TOL_EST * abs(ref2)


# --------------------------------------------------------------------------------- estimators
def test_estimators_match_oracle_with_identical_probes():
    from lip_b200 import lla, stochtrace
    ost, lst, Z, mt, N = _setup("C2_xor")
    D = ost.flat()[0].size
    alpha = 0.9
    cvp_o = O.compute_curvature_approx(ost, Z, mt, alpha, full_set_size=N)
    cvp = lla.compute_curvature_approx(lst, cu(Z), mt, alpha, full_set_size=N)
    rng = np.random.default_rng(2)
    eps = rng.choice([-1.0, 1.0], size=(64, D)).astype(np.float32)
    ref = O.stochastic_trace_estimator_mvp(cvp_o, eps.astype(np.float64))
    got = float(stochtrace.stochastic_trace_estimator_mvp(cvp, D, 0, eps=cu(eps)))
    assert abs(got - ref) <= TOL_EST * abs(ref)
    ref2 = O.hutchpp_v2(cvp_o, eps.astype(np.float64), s1=48, s2=16)
    got2 = float(stochtrace.hutchpp_v2(cvp, lambda _: cu(eps), s1=48, s2=16))
    assert abs(got2 - ref2) <= TOL_EST * abs(ref2)
    g = rng.standard_normal((40, D)).astype(np.float32)
    ref3 = O.hutchpp(cvp_o, g.astype(np.float64))
    got3 = float(stochtrace.hutchpp(cvp, lambda _: cu(g)))
    assert abs(got3 - ref3) <= TOL_EST * abs(ref3)


def test_G3_golden_traces_on_gpu():
    """tests/fixtures.py:201-209, tests/test_stochtrace.py:90-97: hutchpp_v2 exact when s1 >= n."""
    from lip_b200 import stochtrace
    A = np.array([[1.0, 4, 50], [-30, 4.0, 16], [12, 6, 5.0]])
    rng = np.random.default_rng(3)
    for Mx, tr in ((np.diag([1.0, 2.0, 3.0]), 6.0), (A @ A.T, 3894.0)):
        Mt = cu(Mx)
        eps = cu(rng.choice([-1.0, 1.0], size=(8, 3)))
        est = float(stochtrace.hutchpp_v2(lambda v: Mt @ v, lambda _: eps, s1=4, s2=4))
        assert abs(est - tr) <= 2e-5 * tr
        est_d = float(stochtrace.stochastic_trace_estimator_dense(Mt, 0, eps=eps))
        assert abs(est_d - O.stochastic_trace_estimator_dense(Mx, eps.cpu().numpy().astype(np.float64))) <= 1e-4 * tr
    n = 150
    G = rng.standard_normal((n, n))
    M3 = G @ G.T
    M3t = cu(M3)
    g = rng.standard_normal((60, n)).astype(np.float32)
    mat = lambda Mx: M3t @ Mx
    assert abs(float(stochtrace.hutchpp_mvp(mat, n, 0, eps=cu(g))) - O.hutchpp_mvp(lambda Mx: M3 @ Mx, g.astype(np.float64))) \
        <= 1e-4 * np.trace(M3)
    r = rng.choice([-1.0, 1.0], size=(80, n)).astype(np.float32)
    assert abs(float(stochtrace.na_hutchpp_mvp(mat, n, 0, eps=cu(r))) - O.na_hutchpp_mvp(lambda Mx: M3 @ Mx, r.astype(np.float64))) \
        <= 5e-4 * np.trace(M3)
    X = A @ A.T
    Xt = cu(X)
    g3 = rng.standard_normal((20, 3)).astype(np.float32)
    ref = O.hutchpp_inv_mvp(lambda v: X @ v, g3.astype(np.float64))
    got = float(stochtrace.hutchpp_inv_mvp(lambda v: Xt @ v, 3, 0, eps=cu(g3)))
    assert abs(got - ref) <= 2e-3 * abs(ref)


# --------------------------------------------------------------------------------- sampler + predictive
def test_sampler_and_predictive_match_oracle():
    from lip_b200 import lla, sample as S
    ost, lst, Z, mt, N = _setup("C2_xor")
    Z = Z[:12]
    D = ost.flat()[0].size
    alpha = 2.5
    rng = np.random.default_rng(4)
    Eps = rng.standard_normal((3, D)).astype(np.float32)
    ref = O.sample(ost, Z, D, alpha, Eps.astype(np.float64), mt, full_set_size=N)
    got = S.sample(lst, cu(Z), D, alpha, 0, mt, num_samples=3, full_set_size=N, eps=cu(Eps)).cpu().numpy()
    assert got.shape == (3, D)
    assert rel_err(got, ref) < TOL_EST
    Xnew = rng.standard_normal((9, 2)).astype(np.float32)
    refp = O.predict_lla_scalable(ost, Xnew, Z, mt, alpha, Eps.astype(np.float64), full_set_size=N)
    gotp = lla.predict_lla_scalable(lst, cu(Xnew), cu(Z), mt, alpha, full_set_size=N, num_samples=3, eps=cu(Eps)).cpu().numpy()
    assert gotp.shape == (3, 9, 2)
    assert rel_err(gotp, refp) < TOL_EST


def test_unsupported_models_fail_loudly():
    from lip_b200 import ggn, scalemodels
    st = scalemodels.TrainState(params={"Conv_0": {"kernel": np.zeros((5, 5, 1, 6), np.float32)}},
                                apply_fn=scalemodels.LeNet5().apply)
    with pytest.raises(ValueError):              # a LeNet5 whose parameter tree does not have the LeNet geometry
        ggn.compute_ggn_vp(st, torch.zeros(2, 28, 28, 1, device="cuda"), "classifier")
    st3 = scalemodels.TrainState(params={"Conv_0": {"kernel": np.zeros((3, 3, 3, 32), np.float32)}},
                                 apply_fn=scalemodels.ResNet1M().apply)
    with pytest.raises(ValueError):              # a ResNet1M tree without its BatchNorm / blocks / head
        ggn.compute_ggn_vp(st3, torch.zeros(2, 32, 32, 3, device="cuda"), "classifier")

    class Weird:
        def apply(self, *a, **k):
            return None

    st2 = scalemodels.TrainState(params={"Dense_0": {"bias": np.zeros(2, np.float32), "kernel": np.zeros((3, 2), np.float32)}},
                                 apply_fn=Weird().apply)
    with pytest.raises(ValueError):
        ggn.compute_ggn_vp(st2, torch.zeros(2, 3, device="cuda"), "classifier")


# --------------------------------------------------------------------------------- tensor-core path
TC_CONFIGS = {
    # name: (hidden, n_out, in_dim, M, N)
    "tc_small": ([128, 64], 10, 96, 128, 5000),
    "tc_ragged": ([200, 72, 136], 7, 100, 150, 900),
    # fewer bound points than one 128-row MMA tile (the reference's mlp_mnist.yml trains m = 50 inducing points)
    "tc_m50": ([256, 128, 64], 10, 784, 50, 60000),
    "tc_m17": ([136, 72], 5, 100, 17, 300),
}


@pytest.mark.parametrize("name", list(TC_CONFIGS))
def test_tensor_core_path_matches_oracle(name):
    """The tcgen05 3xTF32 path (layers with in,out >= 64 and M >= 16) against the float64 oracle and against the SIMT path."""
    from lip_b200 import ggn, lla
    hidden, n_out, in_dim, M, N = TC_CONFIGS[name]
    ost, lst = make_pair("large", hidden=hidden, n_out=n_out, in_dim=in_dim, seed=77)
    rng = np.random.default_rng(78)
    Z = rng.random((M, in_dim)).astype(np.float32)
    D = ost.flat()[0].size
    V = rng.choice([-1.0, 1.0], size=(5, D)).astype(np.float32)
    V[3:] = rng.standard_normal((2, D)).astype(np.float32)
    alpha = 0.01
    ref_vp = O.compute_curvature_approx(ost, Z, "classifier", alpha, full_set_size=N)
    ref = np.stack([ref_vp(v) for v in V.astype(np.float64)])
    cvp = lla.compute_curvature_approx(lst, cu(Z), "classifier", alpha, full_set_size=N, tensor_path=True)
    assert "tcgen05" in cvp._lip_model.path_name()
    got = cvp(cu(V)).cpu().numpy()
    simt = lla.compute_curvature_approx(lst, cu(Z), "classifier", alpha, full_set_size=N, tensor_path=False)(cu(V)).cpu().numpy()
    e_tc, e_simt = rel_err(got, ref), rel_err(simt, ref)
    print(f"{name}: tc rel err {e_tc:.2e}, simt rel err {e_simt:.2e}")
    assert e_simt < TOL_GGN
    assert e_tc < TOL_GGN
    Wo, WTo = O.compute_W_vps(ost, Z, "classifier", full_set_size=N)
    Wg, WTg = ggn.compute_W_vps(lst, cu(Z), "classifier", full_set_size=N, tensor_path=True)
    ref_wt = np.stack([WTo(v) for v in V.astype(np.float64)])
    assert rel_err(WTg(cu(V)).cpu().numpy(), ref_wt) < TOL_GGN
    U = rng.standard_normal(ref_wt.shape).astype(np.float32)
    assert rel_err(Wg(cu(U)).cpu().numpy(), np.stack([Wo(u) for u in U.astype(np.float64)])) < TOL_GGN


def test_tensor_core_gemm_selftest():
    import ctypes as C
    from lip_b200 import _cabi
    L = _cabi.lib()
    for variant in (0, 1, 2):
        for (M, N, K, b) in [(128, 128, 32, 1), (256, 256, 128, 3), (100, 96, 72, 2), (512, 256, 1024, 2)]:
            err = C.c_float(-1)
            _cabi.check(L.lip_selftest_tc_gemm(variant, M, N, K, b, C.byref(err), None), "selftest")
            assert err.value < 4e-6, (variant, M, N, K, b, err.value)


def test_headline_config_matches_oracle():
    """C3b (MNIST MLP 784-1024-512-256-128-10, M=512) at full size: curvature_vp for 2 probes vs the float64 oracle."""
    from lip_b200 import lla
    ost, lst = make_pair("large", hidden=[1024, 512, 256, 128], n_out=10, in_dim=784, seed=1003, in_shape=(28, 28, 1))
    rng = np.random.default_rng(1004)
    Z = rng.random((512, 784)).astype(np.float32)
    D = ost.flat()[0].size
    assert D == 1494154
    V = rng.choice([-1.0, 1.0], size=(2, D)).astype(np.float32)
    V[1] = rng.standard_normal(D).astype(np.float32)
    ref_vp = O.compute_curvature_approx(ost, Z, "classifier", 1e-3, full_set_size=60000)
    ref = np.stack([ref_vp(v) for v in V.astype(np.float64)])
    for tp in (True, False):
        cvp = lla.compute_curvature_approx(lst, cu(Z), "classifier", 1e-3, full_set_size=60000, tensor_path=tp)
        got = cvp(cu(V)).cpu().numpy()
        errs = [rel_err(got[i], ref[i]) for i in range(2)]
        print(f"C3b tensor_path={tp} ({cvp._lip_model.path_name()}): rel err {errs}")
        assert max(errs) < TOL_GGN


# --------------------------------------------------------------------------------- conv stage programs (LeNet5)
@pytest.mark.parametrize("M", [5, 18])
def test_lenet5_operators_match_oracle(M):
    """M3 (scalemodels.py:11-49): forward, GGN-vector product, W, W^T and the W W^T == GGN identity against the float64
    oracle (torch autograd through the restated LeNet5) on identical weights / points / probes.  M = 5 / 18 leave ragged
    image groups in the fused stage kernels (8 / 4 images per CTA, 2 in flight; lip_cnn_fused.cu)."""
    from lip_b200 import ggn, lla
    ost, lst = make_pair("lenet5", seed=31)
    rng = np.random.default_rng(32)
    N = 60000
    Z = rng.random((M, 28, 28, 1)).astype(np.float32)
    D = ost.flat()[0].size
    assert D == 61706
    bm = ggn._bind(lst, cu(Z), "classifier")
    assert rel_err(bm.outputs().cpu().numpy(), O.model_outputs(ost, Z)) < 2e-6
    V = rng.choice([-1.0, 1.0], size=(3, D)).astype(np.float32)
    V[2] = rng.standard_normal(D).astype(np.float32)
    ref_vp = O.compute_ggn_vp(ost, Z, "classifier", full_set_size=N)
    ref = np.stack([ref_vp(v) for v in V.astype(np.float64)])
    vp = ggn.compute_ggn_vp(lst, cu(Z), "classifier", full_set_size=N)
    got = vp(cu(V)).cpu().numpy()
    assert rel_err(got, ref) < TOL_GGN
    alpha = 5e-3
    cvp = lla.compute_curvature_approx(lst, cu(Z), "classifier", alpha, full_set_size=N)
    assert rel_err(cvp(cu(V[0])).cpu().numpy(), ref[0] + alpha * V[0]) < TOL_GGN
    Wo, WTo = O.compute_W_vps(ost, Z, "classifier", full_set_size=N)
    Wg, WTg = ggn.compute_W_vps(lst, cu(Z), "classifier", full_set_size=N)
    ref_wt = np.stack([WTo(v) for v in V.astype(np.float64)])
    assert rel_err(WTg(cu(V)).cpu().numpy(), ref_wt) < TOL_GGN
    U = rng.standard_normal(ref_wt.shape).astype(np.float32)
    ref_w = np.stack([Wo(u) for u in U.astype(np.float64)])
    assert rel_err(Wg(cu(U)).cpu().numpy(), ref_w) < TOL_GGN
    assert rel_err(Wg(WTg(cu(V))).cpu().numpy(), got) < 2e-5


def test_lenet5_hutchinson_and_predictive():
    """Estimators on top of the conv path: Hutchinson trace with identical probes, batched predictive JVP (lla.py:153)."""
    from lip_b200 import _cabi, ggn, lla, stochtrace
    ost, lst = make_pair("lenet5", seed=33)
    rng = np.random.default_rng(34)
    Z = rng.random((4, 28, 28, 1)).astype(np.float32)
    D = ost.flat()[0].size
    alpha = 5e-3
    eps = rng.choice([-1.0, 1.0], size=(6, D)).astype(np.float32)
    cvp_o = O.compute_curvature_approx(ost, Z, "classifier", alpha, full_set_size=1000)
    ref = O.stochastic_trace_estimator_mvp(cvp_o, eps.astype(np.float64))
    cvp = lla.compute_curvature_approx(lst, cu(Z), "classifier", alpha, full_set_size=1000)
    got = float(stochtrace.stochastic_trace_estimator_mvp(cvp, D, 0, eps=cu(eps)))
    assert abs(got - ref) <= TOL_EST * abs(ref)
    # probes drawn by the estimator itself live in padded rows (stride pad4(D) != D): conv programs repack them
    own = float(stochtrace.stochastic_trace_estimator_mvp(cvp, D, 3, num_samples=8))
    assert np.isfinite(own) and abs(own - ref) < 0.5 * abs(ref)
    # plain Jacobian-vector products at new inputs (factor NONE): J_X w
    Xnew = rng.random((3, 28, 28, 1)).astype(np.float32)
    w = rng.standard_normal((2, D)).astype(np.float32)
    bx = ggn._bind(lst, cu(Xnew), "classifier")
    got_j = bx.wt(cu(w), scale=1.0, factor=_cabi.FACTOR_NONE).cpu().numpy()
    ref_j = np.stack([O.jvp_outputs(ost, Xnew, wi.astype(np.float64)) for wi in w])
    assert rel_err(got_j, ref_j) < TOL_GGN


# --------------------------------------------------------------------------------- residual conv programs (ResNet1M)
def test_resnet1m_operators_match_oracle():
    """M4 (scalemodels.py:70-157): stem + 9 BasicBlocks (two of them strided with a 1x1 shortcut) + global mean + Dense,
    BatchNorm in eval mode with its scale / bias inside the flat parameter vector: forward, GGN-vector product, W^T, W
    against the float64 oracle (torch autograd through the restated network)."""
    from lip_b200 import ggn, lla
    ost, lst = make_pair("resnet1m", n_out=10, seed=41, in_shape=(32, 32, 3))
    rng = np.random.default_rng(42)
    M, N = 2, 49000
    Z = rng.random((M, 32, 32, 3)).astype(np.float32)
    D = ost.flat()[0].size
    assert D == 1084586
    bm = ggn._bind(lst, cu(Z), "classifier")
    assert bm.tensor_layers() == 20          # every conv unit but the 3-channel stem runs on the tcgen05 implicit GEMM
    assert rel_err(bm.outputs().cpu().numpy(), O.model_outputs(ost, Z)) < 5e-6
    V = rng.choice([-1.0, 1.0], size=(2, D)).astype(np.float32)
    V[1] = rng.standard_normal(D).astype(np.float32)
    ref_vp = O.compute_ggn_vp(ost, Z, "classifier", full_set_size=N)
    ref = np.stack([ref_vp(v) for v in V.astype(np.float64)])
    vp = ggn.compute_ggn_vp(lst, cu(Z), "classifier", full_set_size=N)
    got = vp(cu(V)).cpu().numpy()
    assert rel_err(got, ref) < TOL_GGN
    alpha = 5e-3
    cvp = lla.compute_curvature_approx(lst, cu(Z), "classifier", alpha, full_set_size=N)
    assert rel_err(cvp(cu(V[0])).cpu().numpy(), ref[0] + alpha * V[0]) < TOL_GGN
    Wo, WTo = O.compute_W_vps(ost, Z, "classifier", full_set_size=N)
    Wg, WTg = ggn.compute_W_vps(lst, cu(Z), "classifier", full_set_size=N)
    ref_wt = np.stack([WTo(v) for v in V.astype(np.float64)])
    assert rel_err(WTg(cu(V)).cpu().numpy(), ref_wt) < TOL_GGN
    U = rng.standard_normal(ref_wt.shape).astype(np.float32)
    ref_w = np.stack([Wo(u) for u in U.astype(np.float64)])
    assert rel_err(Wg(cu(U)).cpu().numpy(), ref_w) < TOL_GGN


def test_resnet1m_tensor_core_convs_match_simt_and_oracle():
    """The convs with >= 32 channels (stride 1 and 2) run as tcgen05 implicit GEMMs (lip_conv_tc.cu): same operators as the fp32
    SIMT implicit GEMM (tensor_path=False), on a point count that leaves ragged 128-pixel tiles (M = 3 at 8x8 = 1.5 tiles) and
    with Gaussian as well as exactly-TF32 (Rademacher) probes.  Both paths share the bind-time fp32 forward pass, hence the
    ReLU masks; the float64-oracle parity of the tensor path is test_resnet1m_operators_match_oracle (a piecewise-linear
    network at a seed where no pre-activation sits within fp32 rounding of a ReLU kink; at this test's seed one does, and
    the oracle itself moves by 7e-5 under a 1e-7 perturbation of the inputs)."""
    from lip_b200 import ggn
    ost, lst = make_pair("resnet1m", n_out=10, seed=45, in_shape=(32, 32, 3))
    rng = np.random.default_rng(46)
    M, N = 3, 49000
    Z = rng.random((M, 32, 32, 3)).astype(np.float32)
    D = ost.flat()[0].size
    V = rng.choice([-1.0, 1.0], size=(3, D)).astype(np.float32)
    V[1] = rng.standard_normal(D).astype(np.float32)
    V[2, : D // 2] = 0.0
    tc = ggn.compute_ggn_vp(lst, cu(Z), "classifier", full_set_size=N, tensor_path=True)
    assert tc._lip_model.tensor_layers() >= 14, tc._lip_model.path_name()
    simt = ggn.compute_ggn_vp(lst, cu(Z), "classifier", full_set_size=N, tensor_path=False)
    assert simt._lip_model.tensor_layers() == 0
    got_tc, got_simt = tc(cu(V)).cpu().numpy(), simt(cu(V)).cpu().numpy()
    for b in range(3):
        assert rel_err(got_tc[b], got_simt[b]) < TOL_GGN, b
    Wg, WTg = ggn.compute_W_vps(lst, cu(Z), "classifier", full_set_size=N, tensor_path=True)
    Ws, WTs = ggn.compute_W_vps(lst, cu(Z), "classifier", full_set_size=N, tensor_path=False)
    wt = WTg(cu(V))
    assert rel_err(wt.cpu().numpy(), WTs(cu(V)).cpu().numpy()) < TOL_GGN
    U = rng.standard_normal(tuple(wt.shape)).astype(np.float32)
    w_tc = Wg(cu(U)).cpu().numpy()
    assert rel_err(w_tc, Ws(cu(U)).cpu().numpy()) < TOL_GGN


def test_conv_tensor_core_selftest():
    """lip_selftest_conv_tc: the three tcgen05 implicit-GEMM conv roles against the SIMT implicit GEMM on random data."""
    import ctypes as C
    from lip_b200 import _cabi
    L = _cabi.lib()
    # (images, H, W, cin, cout, kernel size, stride, probes)
    shapes = [(4, 32, 32, 32, 32, 3, 1, 2), (3, 16, 16, 64, 64, 3, 1, 2), (5, 8, 8, 128, 128, 3, 1, 3), (4, 16, 16, 32, 64, 1, 1, 2),
              (2, 32, 32, 64, 32, 3, 1, 1), (7, 8, 8, 32, 32, 3, 1, 2), (40, 16, 16, 64, 64, 3, 1, 4),
              (3, 32, 32, 32, 64, 3, 2, 2), (5, 16, 16, 64, 128, 3, 2, 2), (3, 32, 32, 32, 64, 1, 2, 2), (5, 16, 16, 64, 128, 1, 2, 3)]
    for role in (0, 1, 2, 3):
        for (n, H, W, ci, co, k, sd, b) in shapes:
            err, t1, t2 = C.c_float(-1), C.c_float(0), C.c_float(0)
            _cabi.check(L.lip_selftest_conv_tc(role, n, H, W, ci, co, k, sd, b, 0, C.byref(err), C.byref(t1), C.byref(t2), None),
                        "conv selftest")
            assert err.value < 5e-6, (role, n, H, W, ci, co, k, sd, b, err.value)


def test_resnet1m_grayscale_inputs_are_tiled():
    """scalemodels.py:126-127: a 1-channel input is tiled to 3 channels before the stem."""
    from lip_b200 import ggn
    ost, lst = make_pair("resnet1m", n_out=10, seed=43, in_shape=(28, 28, 1))
    rng = np.random.default_rng(44)
    Z = rng.random((2, 28, 28, 1)).astype(np.float32)
    bm = ggn._bind(lst, cu(Z), "classifier")
    assert rel_err(bm.outputs().cpu().numpy(), O.model_outputs(ost, Z)) < 5e-6


# --------------------------------------------------------------------------------- full-size properties (C3b)
def test_headline_size_properties():
    """BASELINE's headline shape (784-1024-512-256-128-10, D = 1,494,154, M = 512) through size-independent properties
    of the operators (the float64 oracle is only run on 2 probes at this size, in test_headline_config_matches_oracle):
    symmetry u.(G v) == v.(G u), linearity, W W^T == GGN, positive semi-definiteness, and SIMT == tcgen05."""
    from lip_b200 import ggn, lla
    ost, lst = make_pair("large", hidden=[1024, 512, 256, 128], n_out=10, in_dim=784, seed=1003, in_shape=(28, 28, 1))
    rng = np.random.default_rng(77)
    Z = cu(rng.random((512, 784)).astype(np.float32))
    D = ost.flat()[0].size
    assert D == 1494154
    N = 60000
    G = ggn.compute_ggn_vp(lst, Z, "classifier", full_set_size=N)
    U = cu(rng.standard_normal((3, D)).astype(np.float32))
    V = cu(rng.choice([-1.0, 1.0], size=(3, D)).astype(np.float32))
    GU, GV = G(U), G(V)
    uGv = (U.double() * GV.double()).sum(1)
    vGu = (V.double() * GU.double()).sum(1)
    assert torch.all((uGv - vGu).abs() <= 2e-5 * (U.double().norm(dim=1) * GV.double().norm(dim=1)))   # symmetry
    assert torch.all((V.double() * GV.double()).sum(1) >= 0) and torch.all((U.double() * GU.double()).sum(1) >= 0)   # PSD
    lin = G(2.0 * U - 0.5 * V)
    assert rel_err(lin.cpu().numpy(), (2.0 * GU - 0.5 * GV).cpu().numpy()) < 1e-5                      # linearity
    Wf, WTf = ggn.compute_W_vps(lst, Z, "classifier", full_set_size=N)
    assert rel_err(Wf(WTf(V)).cpu().numpy(), GV.cpu().numpy()) < 2e-5                                  # W W^T == GGN
    alpha = 1e-3
    S = lla.compute_curvature_approx(lst, Z, "classifier", alpha, full_set_size=N)
    assert rel_err(S(V).cpu().numpy(), (GV + alpha * V).cpu().numpy()) < 1e-5                          # curvature = GGN + alpha I
    G_simt = ggn.compute_ggn_vp(lst, Z, "classifier", full_set_size=N, tensor_path=False)
    assert rel_err(GV.cpu().numpy(), G_simt(V).cpu().numpy()) < 1e-5                                   # both arithmetic paths
    assert rel_err(G(V[0]).cpu().numpy(), GV[0].cpu().numpy()) < 1e-6                                  # un-batched == batched


def test_edge_shapes():
    """One point, one probe, one hidden unit; K = 1 classifier head is rejected for regressors only when K != 1."""
    from lip_b200 import ggn
    ost, lst = make_pair("classifier", hidden=[1], n_out=3, in_dim=1, seed=9)
    Z = np.array([[0.3]], dtype=np.float32)
    D = ost.flat()[0].size
    v = np.arange(1, D + 1, dtype=np.float32) / D
    ref = O.compute_ggn_vp(ost, Z, "classifier", full_set_size=7)(v.astype(np.float64))
    got = ggn.compute_ggn_vp(lst, cu(Z), "classifier", full_set_size=7)(cu(v)).cpu().numpy()
    assert got.shape == (D,) and rel_err(got, ref) < TOL_GGN
    with pytest.raises(ValueError):
        ggn.compute_ggn_vp(lst, cu(np.zeros((2, 5), np.float32)), "classifier")          # wrong feature count
    with pytest.raises(ValueError):
        ggn.compute_ggn_vp(lst, cu(Z), "classifier")(cu(np.zeros(D + 1, np.float32)))    # wrong vector length


def test_packed_rademacher_probes_round_trip():
    """Bit-exact: pack (host, numpy.packbits) -> unpack (device) reproduces the +-1 probe matrix, ragged n included."""
    from lip_b200 import stochtrace
    rng = np.random.default_rng(8)
    for B, n in ((1, 1), (3, 8), (5, 1003), (2, 61706)):
        eps = rng.choice([-1.0, 1.0], size=(B, n)).astype(np.float32)
        bits = stochtrace.pack_rademacher(eps)
        assert bits.shape == (B, (n + 7) // 8) and bits.dtype == np.uint8
        got = stochtrace.unpack_rademacher(torch.as_tensor(bits, device="cuda"), n).cpu().numpy()
        np.testing.assert_array_equal(got, eps)
    with pytest.raises(ValueError):
        stochtrace.unpack_rademacher(torch.zeros(2, 3), 24)


def test_exact_tf32_probes_take_the_zero_lo_path():
    """+-1 (Rademacher) and one-hot probe blocks are exactly TF32: the split kernel finds no lo part and the tcgen05 GEMMs
    skip that operand's lo loads / MMAs.  The result must equal the oracle exactly as for general probes, and a batch that
    mixes exact and general rows must take the general path."""
    from lip_b200 import ggn
    hidden, n_out, in_dim, M, N = TC_CONFIGS["tc_ragged"]
    ost, lst = make_pair("large", hidden=hidden, n_out=n_out, in_dim=in_dim, seed=300)
    rng = np.random.default_rng(301)
    Z = rng.standard_normal((M, in_dim)).astype(np.float32)
    D = ost.flat()[0].size
    ref_vp = O.compute_ggn_vp(ost, Z, "classifier", full_set_size=N)
    vp = ggn.compute_ggn_vp(lst, cu(Z), "classifier", full_set_size=N, tensor_path=True)
    V_pm = rng.choice([-1.0, 1.0], size=(3, D)).astype(np.float32)
    V_hot = np.zeros((2, D), np.float32); V_hot[0, 5] = 1.0; V_hot[1, D - 3] = -2.0
    V_mix = V_pm.copy(); V_mix[1] = rng.standard_normal(D).astype(np.float32)
    for V in (V_pm, V_hot, V_mix):
        ref = np.stack([ref_vp(v) for v in V.astype(np.float64)])
        assert rel_err(vp(cu(V)).cpu().numpy(), ref) < TOL_GGN


def test_ggn_vp_is_cuda_graph_capturable():
    """The hot call enqueues asynchronously on the caller's stream with no allocation / host sync inside the library
    (include/lip_b200.h contract): one lip_ggn_vp call captured into a CUDA graph replays to the same result, on both
    arithmetic paths, with new probe values written into the captured input buffer."""
    from lip_b200 import lla
    hidden, n_out, in_dim, M, N = TC_CONFIGS["tc_small"]
    ost, lst = make_pair("large", hidden=hidden, n_out=n_out, in_dim=in_dim, seed=310)
    rng = np.random.default_rng(311)
    Z = cu(rng.standard_normal((M, in_dim)).astype(np.float32))
    D = ost.flat()[0].size
    for tp in (False, True):
        cvp = lla.compute_curvature_approx(lst, Z, "classifier", 0.25, full_set_size=N, tensor_path=tp)
        V = cu(rng.standard_normal((6, D)).astype(np.float32))
        eager = cvp(V).clone()
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            cvp(V)                                   # warm-up on the capture stream (workspace growth, attribute setup)
        torch.cuda.current_stream().wait_stream(side)
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph):
            out = cvp(V)
        graph.replay()
        torch.cuda.synchronize()
        assert rel_err(out.cpu().numpy(), eager.cpu().numpy()) < 1e-6
        V2 = cu(rng.choice([-1.0, 1.0], size=(6, D)).astype(np.float32))
        expect = cvp(V2).clone()
        V.copy_(V2)                                  # new probes into the captured input buffer
        graph.replay()
        torch.cuda.synchronize()
        assert rel_err(out.cpu().numpy(), expect.cpu().numpy()) < 1e-6


# --------------------------------------------------------------------------------- the production caller
@pytest.mark.parametrize("kind", ["classifier", "regressor"])
def test_scalable_kl_objective_forward_matches_oracle(kind):
    """train_inducing.py:87-173 (forward value): Hutch++ v2 trace of S_X S_Z^{-1} (Woodbury) + GKL logdet, with identical
    probes, against the float64 restatement."""
    from lip_b200 import train_inducing
    if kind == "classifier":
        ost, lst = make_pair("classifier", hidden=[16, 16], n_out=2, in_dim=2, seed=50)
        M, k = 32, None
    else:
        ost, lst = make_pair("regressor", hidden=[8, 8], n_out=1, in_dim=1, seed=51, logvar=0.3)
        M, k = 30, 20
    in_dim = 2 if kind == "classifier" else 1
    rng = np.random.default_rng(52)
    Z = rng.standard_normal((M, in_dim)).astype(np.float32)
    X = rng.standard_normal((48, in_dim)).astype(np.float32)
    D = ost.flat()[0].size
    probes = rng.choice([-1.0, 1.0], size=(40, D)).astype(np.float32)
    alpha, N = 0.9, 800
    ref = O.alternative_objective_scalable(Z, X, ost, alpha, kind, probes.astype(np.float64), full_set_size=N, slq_samples=2,
                                           slq_num_matvecs=k)
    got = float(train_inducing.alternative_objective_scalable(cu(Z), cu(X), lst, alpha, kind, 0, full_set_size=N, st_samples=40,
                                                              slq_samples=2, slq_num_matvecs=k, probes=cu(probes)))
    assert abs(got - ref) <= TOL_EST * abs(ref), (got, ref)


def test_materialize_covariance_probes_an_operator():
    """lla.py:160-217: diagonal / full matrix of a linear operator by unit-vector probing (float64 buffers)."""
    from lip_b200 import lla
    rng = np.random.default_rng(60)
    N, out_dim = 5, 3
    A = rng.standard_normal((N * out_dim, N * out_dim))
    Cov = A @ A.T
    Ct = torch.as_tensor(Cov, device="cuda", dtype=torch.float64)
    f = lambda e: (Ct @ e.reshape(-1).double()).reshape(N, out_dim)
    diag = lla.materialize_covariance(f, N, out_dim, mode="diag").cpu().numpy()
    np.testing.assert_allclose(diag, np.diag(Cov).reshape(N, out_dim), rtol=1e-12)
    full = lla.materialize_covariance(f, N, out_dim, mode="full").cpu().numpy()
    np.testing.assert_allclose(full, Cov, rtol=1e-12)
    ref_diag = O.materialize_covariance(lambda e: Cov @ np.asarray(e).reshape(-1), N, out_dim, mode="diag")
    np.testing.assert_allclose(diag, np.asarray(ref_diag).reshape(N, out_dim), rtol=1e-12)
    with pytest.raises(ValueError):
        lla.materialize_covariance(f, N, out_dim, mode="banana")


def test_exact_tf32_probes_are_read_in_place():
    """lip_ggn_vp_ex with LIP_PROBES_EXACT_TF32 (round 2): +-1 Rademacher probes handed over in padded rows (stochtrace._rademacher,
    unpack_rademacher) are the tcgen05 JVP GEMMs' B operand where they lie - no TF32 split pass.  Same numbers as the split path
    (bit for bit: the hi part IS the probe, the lo part is zero) and as the float64 oracle; rows that are not TMA-addressable, and
    unmarked data, take the split path."""
    from lip_b200 import lla, stochtrace, _cabi
    from lip_b200._runtime import exact_tf32_block
    L = _cabi.lib()
    for name in ("tc_ragged", "tc_small"):            # D = 45,159 (D % 4 = 3) and D = 21,322 (D % 4 = 2)
        hidden, n_out, in_dim, M, N = TC_CONFIGS[name]
        ost, lst = make_pair("large", hidden=hidden, n_out=n_out, in_dim=in_dim, seed=77)
        rng = np.random.default_rng(81)
        Z = rng.random((M, in_dim)).astype(np.float32)
        D = ost.flat()[0].size
        E = rng.choice([-1.0, 1.0], size=(5, D)).astype(np.float32)
        cvp = lla.compute_curvature_approx(lst, cu(Z), "classifier", 0.02, full_set_size=N, tensor_path=True)
        ref_vp = O.compute_curvature_approx(ost, Z, "classifier", 0.02, full_set_size=N)
        ref = np.stack([ref_vp(v) for v in E.astype(np.float64)])
        plain = cvp(cu(E))                                        # unmarked, contiguous [B, D]: split path
        Vp = exact_tf32_block(5, D)
        Vp.copy_(cu(E))
        assert Vp.stride(0) % 4 == 0 and Vp.stride(0) >= D
        l0 = L.lip_launch_count()
        fast = cvp(Vp)
        n_fast = L.lip_launch_count() - l0
        l0 = L.lip_launch_count()
        cvp(cu(E))
        n_plain = L.lip_launch_count() - l0
        assert n_fast < n_plain                                    # the split kernels are gone
        assert torch.equal(fast, plain)
        assert rel_err(fast.cpu().numpy(), ref) < TOL_GGN
        # the estimator entry point generates such probes itself
        tr = float(stochtrace.stochastic_trace_estimator_mvp(cvp, D, 3, num_samples=16))
        eps = stochtrace._rademacher(3, (16, D))
        assert stochtrace.is_exact_tf32(eps) and eps.stride(0) % 4 == 0
        tr_ref = O.stochastic_trace_estimator_mvp(ref_vp, eps.cpu().numpy().astype(np.float64))
        assert abs(tr - tr_ref) <= TOL_EST * abs(tr_ref)
        # packed wire format -> padded rows on the device
        bits = torch.as_tensor(stochtrace.pack_rademacher(E), device="cuda")
        up = stochtrace.unpack_rademacher(bits, D)
        assert stochtrace.is_exact_tf32(up) and torch.equal(up, cu(E)) and torch.equal(cvp(up), plain)
        # a slice of the padded block is a new tensor object: unmarked, still row-strided -> split path with the padded stride
        assert torch.equal(cvp(Vp[1:4]), plain[1:4])
