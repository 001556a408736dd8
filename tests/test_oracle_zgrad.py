"""CPU checks of the oracle's Z-gradients: torch.func.grad over the literal restatement against
central finite differences of the (independently written, numpy-based) forward oracle, the identity between the exact-Gram and
the dense objective (they differ by Z-independent constants, src/Untitled-1.md:1-2), and the committed vectors."""
import importlib.util
import math
import os

import numpy as np

from helpers import make_pair
from oracle import lip_oracle as O

HERE = os.path.dirname(os.path.abspath(__file__))


def _fd(fun, Z, eps=1e-6):
    g = np.zeros_like(Z)
    for idx in np.ndindex(*Z.shape):
        Zp, Zm = Z.copy(), Z.copy()
        Zp[idx] += eps
        Zm[idx] -= eps
        g[idx] = (fun(Zp) - fun(Zm)) / (2 * eps)
    return g


def _small(kind, n_out, logvar=0.0):
    ost, _ = make_pair(kind, hidden=[6, 5], n_out=n_out, in_dim=2, seed=5, logvar=logvar)
    rng = np.random.default_rng(6)
    Z = rng.standard_normal((4, 2))
    D = ost.flat()[0].size
    return ost, Z, D, rng


def test_operator_zgrads_match_finite_differences():
    for kind, n_out, mt, lv in (("classifier", 3, "classifier", 0.0), ("regressor", 1, "regressor", 0.4)):
        ost, Z, D, rng = _small(kind, n_out, lv)
        U, V = rng.standard_normal((2, D)), rng.standard_normal((2, D))
        Y = rng.standard_normal((2, 4, n_out))
        g = O.ggn_vp_zgrad(ost, Z, mt, U, V, full_set_size=40)
        fd = _fd(lambda Zv: sum((u * O.compute_ggn_vp(ost, Zv, mt, full_set_size=40)(v)).sum() for u, v in zip(U, V)), Z)
        assert np.abs(g - fd).max() <= 1e-6 * np.abs(fd).max()
        Wz, WTz = O.W_vps_zgrad(ost, Z, mt, full_set_size=40)
        shp = (4,) if mt == "regressor" else (4, n_out)
        fdW = _fd(lambda Zv: sum((u * O.compute_W_vps(ost, Zv, mt, full_set_size=40)[0](y.reshape(shp))).sum() for u, y in zip(U, Y)), Z)
        assert np.abs(Wz(U, Y) - fdW).max() <= 1e-6 * np.abs(fdW).max()
        fdT = _fd(lambda Zv: sum((y.reshape(shp) * O.compute_W_vps(ost, Zv, mt, full_set_size=40)[1](v)).sum() for y, v in zip(Y, V)), Z)
        assert np.abs(WTz(Y, V) - fdT).max() <= 1e-6 * np.abs(fdT).max()
        Cb = rng.standard_normal((2, 4, n_out))
        fdJ = _fd(lambda Zv: sum((c * O.jvp_outputs(ost, Zv, v)).sum() for c, v in zip(Cb, V)), Z)
        assert np.abs(O.jvp_zgrad(ost, Z, Cb, V) - fdJ).max() <= 1e-6 * np.abs(fdJ).max()


def test_exact_and_dense_objectives_share_their_gradient_and_match_the_forward_oracle():
    ost, Z, D, rng = _small("classifier", 3)
    X = rng.standard_normal((7, 2))
    v_d, g_d = O.variational_grad_dense(Z, X, ost, 0.5, "classifier", full_set_size=70)
    v_e, g_e = O.variational_grad_scalable_exact(Z, X, ost, 0.5, "classifier", full_set_size=70)
    assert np.abs(g_d - g_e).max() <= 1e-10 * np.abs(g_d).max()
    # dense value from the numpy forward oracle (compute_ggn_dense, ggn.py:149-193)
    S = O.compute_ggn_dense(ost, X, "classifier", 70)[0] + 0.5 * np.eye(D)
    Sz = O.compute_ggn_dense(ost, Z, "classifier", 70)[0] + 0.5 * np.eye(D)
    ref = np.trace(S @ np.linalg.inv(Sz)) + np.linalg.slogdet(Sz)[1]
    assert abs(v_d - ref) <= 1e-9 * abs(ref)
    # the exact-Gram form equals the dense form up to the dropped Z-independent constant (train_inducing.py:63 comment):
    # tr(S Sz^-1) + logdet Sz = [D + (gamma/alpha) tr(W^T W)] + exact     (identity of src/Untitled-1.md:1-2)
    Wx, WTx = O.compute_W_vps(ost, X, "classifier")
    G = O.build_WTW(Wx, WTx, (7, 3), 21)
    const = D + (70 / 7) / 0.5 * np.trace(G)
    assert abs((v_e + const) - v_d) <= 1e-8 * abs(v_d)


def test_log_marginal_likelihood_gradient():
    ost, Z, D, rng = _small("classifier", 3)
    X = rng.standard_normal((7, 2))
    v, g = O.log_marginal_likelihood(0.7, X, ost, "classifier", full_set_size=70)
    h = 1e-5
    vp = O.log_marginal_likelihood(0.7 * math.exp(h), X, ost, "classifier", full_set_size=70)[0]
    vm = O.log_marginal_likelihood(0.7 * math.exp(-h), X, ost, "classifier", full_set_size=70)[0]
    assert abs((vp - vm) / (2 * h) - g) <= 1e-6 * abs(g)


def test_oracle_regenerates_committed_zgrad_vectors():
    gold = np.load(os.path.join(HERE, "golden", "zgrad_v1.npz"))
    spec = importlib.util.spec_from_file_location("make_golden_zgrad", os.path.join(HERE, "golden", "make_golden_zgrad.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    fresh = mod.build()
    assert sorted(fresh) == sorted(gold.files)
    for k in gold.files:
        np.testing.assert_allclose(np.asarray(fresh[k]), gold[k], rtol=1e-8, atol=1e-11, err_msg=k)
