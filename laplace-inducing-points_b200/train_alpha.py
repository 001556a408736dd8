"""Drop-in mirror of /root/reference/src/train_alpha.py:13-59 on the B200 path: the Laplace log marginal likelihood as a
function of the prior precision alpha, through the dense Gram W^T W (lip_gram_wtw), and one optimiser step on log(alpha).
SURVEY §8f row f3."""
from __future__ import annotations

import math

import torch

from ._runtime import dev_f32
from .ggn import build_WTW, compute_W_vps
from .utils import flatten_nn_params


def _parts(X, state, model_type, full_set_size):
    Xt = dev_f32(X)
    n = int(Xt.shape[0])
    N = full_set_size or n
    rescale = N / n                                                                         # :22
    flat_p, _ = flatten_nn_params(state.params)                                             # D excludes logvar (:24-26)
    D = int(flat_p.numel())
    W, WT = compute_W_vps(state, Xt, model_type, full_set_size=None)                        # :28
    bm = W._lip_model
    inner_shape = (n,) if model_type == "regressor" else (n, bm.K)
    d = n * bm.K
    WTW = build_WTW(W, WT, inner_shape, d, dtype=torch.float32, block=1).double()           # :32
    lam = torch.linalg.eigvalsh(WTW).clamp_min(0.0)                                         # spectrum of the PSD Gram (library call)
    return rescale, D, lam, float((flat_p.double() @ flat_p.double()).item())


def _lml(alpha, rescale, D, lam, sq):
    logdet_term = torch.log1p(rescale / alpha * lam).sum() + D * math.log(alpha)            # :35-36 slogdet(I + rescale/alpha WTW)
    log_prior = -0.5 * alpha * sq + 0.5 * D * math.log(alpha)                               # :39-42
    return log_prior - 0.5 * logdet_term                                                    # :44


def log_marginal_likelihood(alpha, X, state, model_type, full_set_size=None):
    """train_alpha.py:13-44: log p(D | alpha) up to alpha-independent constants."""
    rescale, D, lam, sq = _parts(X, state, model_type, full_set_size)
    return _lml(float(alpha), rescale, D, lam, sq).float()


def log_marginal_likelihood_value_and_grad(log_alpha, X, state, model_type, full_set_size=None):
    """(L, dL/d log alpha) — what jax.grad(loss_fn)(log_alpha) differentiates in update_alpha (:54-56), in closed form:
    dL/d log a = -a theta.theta / 2 + sum_j (r lam_j / a) / (1 + r lam_j / a) / 2."""
    alpha = math.exp(float(log_alpha))
    rescale, D, lam, sq = _parts(X, state, model_type, full_set_size)
    x = rescale / alpha * lam
    grad = -0.5 * alpha * sq + 0.5 * (x / (1.0 + x)).sum()
    return _lml(alpha, rescale, D, lam, sq).float(), grad.float()


def update_alpha(log_alpha, opt_state, opt, *lm_args):
    """train_alpha.py:47-59: gradient ascent on log alpha, written as an optax-protocol descent step on -L."""
    _, g = log_marginal_likelihood_value_and_grad(log_alpha, *lm_args)
    la = torch.as_tensor(float(log_alpha), dtype=torch.float32, device=g.device)
    updates, new_state = opt.update(-g, opt_state, la)
    return la + updates, new_state
