"""GPU parity for SURVEY §8 rows f1 / f3: the deterministic inducing-point objectives of train_inducing.py (exact-Gram form
:26-84 and dense form :175-192), their gradients with respect to Z (what jax.value_and_grad yields, :194-195), one
optimize_step (:198-232) and the alpha evidence of train_alpha.py:13-59, against the float64 autograd oracle."""
import math

import numpy as np
import pytest
import torch

from helpers import make_pair, rel_err
from oracle import lip_oracle as O
from test_gpu_parity import cu

pytestmark = pytest.mark.gpu

TOL_VALUE = 1e-4      # objective values (north_star: trace / logdet estimates)
TOL_GRAD = 5e-4       # gradients of those values: they pass through (I/beta + G/alpha)^-1 of an fp32 Gram (cond ~ 1 + beta/alpha lmax)

CASES = {
    # name: (kind, hidden, n_out, in_dim, M, n_x, N, alpha, logvar)
    "xor": ("classifier", [16, 16], 2, 2, 12, 20, 800, 0.05, 0.0),
    "cls3": ("classifier", [8, 8], 3, 2, 7, 11, 90, 0.5, 0.0),
    "sine": ("regressor", [8, 8], 1, 1, 10, 16, 240, 0.5, 0.3),
}


def _case(name):
    kind, hidden, n_out, in_dim, M, nx, N, alpha, logvar = CASES[name]
    ost, lst = make_pair(kind, hidden=hidden, n_out=n_out, in_dim=in_dim, seed=31, logvar=logvar)
    rng = np.random.default_rng(32)
    Z = rng.standard_normal((M, in_dim)).astype(np.float32)
    X = rng.standard_normal((nx, in_dim)).astype(np.float32)
    return ost, lst, Z, X, ("regressor" if kind == "regressor" else "classifier"), N, alpha


@pytest.mark.parametrize("name", list(CASES))
def test_exact_objective_value_and_zgrad(name):
    from lip_b200 import train_inducing as TI
    ost, lst, Z, X, mt, N, alpha = _case(name)
    ref_v, ref_g = O.variational_grad_scalable_exact(Z, X, ost, alpha, mt, full_set_size=N)
    val = TI.alternative_objective_scalable_exact(cu(Z), cu(X), lst, alpha, mt, None, full_set_size=N)
    assert abs(float(val) - ref_v) <= TOL_VALUE * abs(ref_v)
    v, g = TI.variational_grad_scalable_exact(cu(Z), cu(X), lst, alpha, mt, None, full_set_size=N)
    assert abs(float(v) - ref_v) <= TOL_VALUE * abs(ref_v)
    assert g.shape == Z.shape and rel_err(g.cpu().numpy(), ref_g) < TOL_GRAD


@pytest.mark.parametrize("name", list(CASES))
def test_dense_objective_value_and_zgrad(name):
    from lip_b200 import train_inducing as TI
    ost, lst, Z, X, mt, N, alpha = _case(name)
    ref_v, ref_g = O.variational_grad_dense(Z, X, ost, alpha, mt, full_set_size=N)
    v, g = TI.variational_grad_dense(cu(Z), cu(X), lst, alpha, mt, None, full_set_size=N)
    assert abs(float(v) - ref_v) <= TOL_VALUE * abs(ref_v)
    assert abs(float(TI.alternative_objective_dense(cu(Z), cu(X), lst, alpha, mt, None, full_set_size=N)) - ref_v) <= TOL_VALUE * abs(ref_v)
    assert rel_err(g.cpu().numpy(), ref_g) < TOL_GRAD


def test_optimize_step_dense_and_exact_agree_with_oracle_gradient_step():
    from lip_b200 import train_inducing as TI, utils
    ost, lst, Z, X, mt, N, alpha = _case("xor")
    _, ref_g = O.variational_grad_dense(Z, X, ost, alpha, mt, full_set_size=N)
    opt = utils.sgd(1e-3)
    for kw in (dict(scalable=False), dict(scalable=True, exact=True)):
        Znew, state, loss = TI.optimize_step(cu(Z), cu(X), lst, alpha, opt.init(cu(Z)), 0, opt, None, mt, full_set_size=N, **kw)
        assert rel_err((Znew.cpu().numpy() - Z), -1e-3 * ref_g) < 5e-4
    adam = utils.adam(1e-2)
    Znew, st, _ = TI.optimize_step(cu(Z), cu(X), lst, alpha, adam.init(cu(Z)), 0, adam, None, mt, full_set_size=N, scalable=False)
    # first Adam step: -lr * sign(g) up to eps
    np.testing.assert_allclose(Znew.cpu().numpy() - Z, -1e-2 * np.sign(ref_g), atol=1e-4)


def test_scalable_gradient_is_exact_and_hutchinson_form_agrees_in_expectation():
    """variational_grad_scalable: the loss is the stochastic estimate, dZ the exact gradient of tr(S_X S_Z^-1) + logdet S_Z (default), which
    must equal the deterministic oracle gradient.  gradient="hutchinson": with the probe set sqrt(D) e_1 .. sqrt(D) e_D the Hutchinson
    average IS the trace, so it must reproduce the same gradient; with Rademacher probes it is checked loosely (statistical)."""
    from lip_b200 import train_inducing as TI
    ost, lst, Z, X, mt, N, alpha = _case("xor")
    _, ref_g = O.variational_grad_dense(Z, X, ost, alpha, mt, full_set_size=N)
    D = ost.flat()[0].size
    probes = np.random.default_rng(33).choice([-1.0, 1.0], size=(64, D)).astype(np.float32)
    loss, g = TI.variational_grad_scalable(cu(Z), cu(X), lst, alpha, mt, 0, full_set_size=N, slq_num_matvecs=8, probes=cu(probes))
    assert g.shape == Z.shape and rel_err(g.cpu().numpy(), ref_g) < TOL_GRAD
    assert math.isfinite(float(loss))
    _, g = TI.variational_grad_scalable(cu(Z), cu(X), lst, alpha, mt, 0, full_set_size=N, slq_num_matvecs=8, probes=cu(probes),
                                        gradient="woodbury")
    assert rel_err(g.cpu().numpy(), ref_g) < 2e-3          # every S_Z^-1 application is an fp32 Woodbury solve
    basis = (math.sqrt(D) * np.eye(D)).astype(np.float32)
    _, g = TI.variational_grad_scalable(cu(Z), cu(X), lst, alpha, mt, 0, full_set_size=N, slq_num_matvecs=8, probes=cu(basis),
                                        gradient="hutchinson")
    assert rel_err(g.cpu().numpy(), ref_g) < 2e-3
    probes = np.random.default_rng(33).choice([-1.0, 1.0], size=(4096, D)).astype(np.float32)
    _, g = TI.variational_grad_scalable(cu(Z), cu(X), lst, alpha, mt, 0, full_set_size=N, slq_num_matvecs=8, probes=cu(probes),
                                        gradient="hutchinson")
    g = g.cpu().numpy()
    cos = float((g * ref_g).sum() / (np.linalg.norm(g) * np.linalg.norm(ref_g)))
    assert cos > 0.9, (cos, rel_err(g, ref_g))
    with pytest.raises(ValueError):
        TI.variational_grad_scalable(cu(Z), cu(X), lst, alpha, mt, 0, full_set_size=N, slq_num_matvecs=8, probes=cu(probes[:32]), gradient="x")


@pytest.mark.parametrize("name", ["xor", "sine"])
def test_log_marginal_likelihood_and_alpha_step(name):
    from lip_b200 import train_alpha as TA, utils
    ost, lst, Z, X, mt, N, alpha = _case(name)
    ref_v, ref_g = O.log_marginal_likelihood(alpha, X, ost, mt, full_set_size=N)
    v = TA.log_marginal_likelihood(alpha, cu(X), lst, mt, full_set_size=N)
    assert abs(float(v) - ref_v) <= TOL_VALUE * abs(ref_v)
    v2, g = TA.log_marginal_likelihood_value_and_grad(math.log(alpha), cu(X), lst, mt, full_set_size=N)
    assert abs(float(g) - ref_g) <= TOL_VALUE * abs(ref_g)
    opt = utils.sgd(0.1)
    la, _ = TA.update_alpha(math.log(alpha), opt.init(0.0), opt, cu(X), lst, mt, N)
    assert abs(float(la) - (math.log(alpha) + 0.1 * ref_g)) < 1e-4 * max(1.0, abs(ref_g))
