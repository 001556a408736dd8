"""Drop-in mirror of the evaluation consumer of the hot path, /root/reference/scale_experiments/evaluate.py:40-231
(SURVEY §8 row f4): Monte-Carlo softmax NLL / accuracy of the linearized-Laplace predictive, Brier score, expected calibration
error, max-probability OOD score and AUROC.  The logit samples come from lla.predict_lla_scalable (one batched JVP for all
samples); the MC-softmax reduction is lip_mc_softmax_predictive (csrc/lip_eval.cu).  Unlike the reference, which rebuilds the
posterior sampler (Gram matrix + factorisation, sample.py:77) for every test batch, eval_dataset builds it once."""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch

from . import _cabi as cabi
from ._runtime import dev_f32, ptr, stream
from .lla import predict_lla_scalable
from .sample import inv_matsqrt_vp
from .utils import flatten_nn_params


def mc_softmax_predictive(logit_samples, y=None):
    """logit_samples [S, B, C] (+ labels [B]) -> (log of the MC-averaged probability of the true class [B] or None, mean probs [B, C])."""
    ls = dev_f32(logit_samples)
    if ls.dim() != 3:
        raise ValueError(f"logit samples must be [S, B, C], got {tuple(ls.shape)}")
    S, B, Cc = (int(v) for v in ls.shape)
    mean = torch.empty(B, Cc, device=ls.device, dtype=torch.float32)
    lab = lavg = None
    if y is not None:
        lab = torch.as_tensor(y).to(device=ls.device).reshape(-1).to(torch.int32).contiguous()     # y.squeeze().astype(int32), :122
        if lab.numel() != B:
            raise ValueError(f"{lab.numel()} labels for {B} examples")
        lavg = torch.empty(B, device=ls.device, dtype=torch.float32)
    cabi.check(cabi.lib().lip_mc_softmax_predictive(ptr(ls), ptr(lab), ptr(lavg), ptr(mean), S, B, Cc, stream()),
               "lip_mc_softmax_predictive")
    return lavg, mean


def batch_nll(state, x, y, Z, *, alpha, full_set_size, model_type, num_mc_samples, rng, scalable=True, return_mean=False,
              sampler=None, eps=None):
    """evaluate.py:98-152 -> (nll, acc[, mean probs]).  scalable=False (the tfp dense predictive, lla.py:42-79) is outside the hot
    path and raises."""
    if not scalable:
        raise ValueError("batch_nll: only the scalable predictive (predict_lla_scalable) is on the B200 path")
    logit_samples = predict_lla_scalable(state, x, Z, model_type=model_type, alpha=alpha, full_set_size=full_set_size,
                                         num_samples=num_mc_samples, key=rng, eps=eps, sampler=sampler)       # (S, B, C)
    log_avg_prob, mean = mc_softmax_predictive(logit_samples, y)
    nll = -log_avg_prob.mean()                                                                                  # :139
    yv = torch.as_tensor(y).to(mean.device).reshape(-1)
    acc = (mean.argmax(-1) == yv).float().mean()                                                                # :146
    if return_mean:
        return nll, acc, mean
    return nll, acc


def _sampler_for(state, Z, alpha, model_type, full_set_size):
    D = int(flatten_nn_params(state.params)[0].numel())
    return inv_matsqrt_vp(state, Z, D, alpha, model_type, full_set_size=full_set_size, key=None)


def _split(rng):
    """jax.random.split stand-in: (carry, sub) integer keys."""
    if isinstance(rng, torch.Generator):
        return rng, rng
    rng = int(rng)
    return (rng * 6364136223846793005 + 1442695040888963407) % (1 << 63), rng


def eval_dataset(state, dataloader, Z, alpha, full_set_size, model_type, num_mc_samples, rng, scalable=True):
    """evaluate.py:155-182 -> (mean NLL, accuracy) over an iterable of (x, y) batches."""
    sampler = _sampler_for(state, Z, alpha, model_type, full_set_size)
    tot_nll, tot_correct, tot_N = 0.0, 0.0, 0
    for x_b, y_b in dataloader:
        rng, sub = _split(rng)
        nll, acc = batch_nll(state, x_b, y_b, Z, alpha=alpha, full_set_size=full_set_size, model_type=model_type,
                             num_mc_samples=num_mc_samples, rng=sub, scalable=scalable, sampler=sampler)
        bs = int(torch.as_tensor(y_b).reshape(-1).shape[0])
        tot_nll += float(nll) * bs
        tot_correct += float(acc) * bs
        tot_N += bs
    return tot_nll / tot_N, tot_correct / tot_N


def eval_dataset_extended(state, dataloader, Z, alpha, full_set_size, model_type, num_mc_samples, rng, scalable=True):
    """evaluate.py:185-231 -> (NLL, accuracy, Brier, ECE, probs, labels)."""
    sampler = _sampler_for(state, Z, alpha, model_type, full_set_size)
    tot_nll, tot_correct, tot_N = 0.0, 0.0, 0
    all_probs, all_labels = [], []
    for x_b, y_b in dataloader:
        rng, sub = _split(rng)
        nll, acc, mean = batch_nll(state, x_b, y_b, Z, alpha=alpha, full_set_size=full_set_size, model_type=model_type,
                                   num_mc_samples=num_mc_samples, rng=sub, scalable=scalable, return_mean=True, sampler=sampler)
        bs = int(mean.shape[0])
        tot_nll += float(nll) * bs
        tot_correct += float(acc) * bs
        tot_N += bs
        all_probs.append(mean.cpu().numpy())
        all_labels.append(np.asarray(torch.as_tensor(y_b).cpu()).reshape(-1))
    probs = np.concatenate(all_probs, axis=0)
    labels = np.concatenate(all_labels, axis=0)
    return tot_nll / tot_N, tot_correct / tot_N, brier_score(probs, labels), ece(probs, labels), probs, labels


# ---- calibration / OOD metrics (host side, as in the reference: numpy on the concatenated mean probabilities) ----
def brier_score(probs, labels) -> float:
    """evaluate.py:40-43: mean over examples of sum_c (p_c - [c == y])^2."""
    probs = np.asarray(probs, dtype=np.float64)
    labels = np.asarray(labels).astype(np.int64).reshape(-1)
    p_true = probs[np.arange(len(labels)), labels]
    return float(np.mean((probs ** 2).sum(axis=1) - 2.0 * p_true + 1.0))


def ece(probs, labels, n_bins: int = 15) -> float:
    """evaluate.py:45-63: histogram ECE over n_bins half-open confidence bins [lo, hi) on linspace(0, 1) edges (a confidence of
    exactly 1.0 falls in no bin, as in the reference)."""
    probs = np.asarray(probs)
    labels = np.asarray(labels).reshape(-1)
    conf = probs.max(axis=1)
    correct = (probs.argmax(axis=1) == labels).astype(np.float64)
    edges = np.linspace(0.0, 1.0, n_bins + 1)
    total = 0.0
    for lo, hi in zip(edges[:-1], edges[1:]):
        sel = (conf >= lo) & (conf < hi)
        if sel.any():
            total += abs(float(conf[sel].mean()) - float(correct[sel].mean())) * float(sel.mean())
    return float(total)


def ood_scores(probs):
    """evaluate.py:65-67: higher = more in-distribution-like."""
    return -np.asarray(probs).max(axis=1)


def auroc_ood(state, id_probs, ood_loader, Z, alpha, full_set_size, model_type, num_mc_samples, rng, scalable=True):
    """evaluate.py:70-93: AUROC of the max-probability score separating in-distribution (label 0) from OOD (label 1) inputs."""
    from sklearn.metrics import roc_auc_score
    sampler = _sampler_for(state, Z, alpha, model_type, full_set_size)
    ood_probs = []
    for xb, _ in ood_loader:
        rng, sub = _split(rng)
        logits = predict_lla_scalable(state, xb, Z, model_type=model_type, alpha=alpha, full_set_size=full_set_size,
                                      num_samples=num_mc_samples, key=sub, sampler=sampler)
        ood_probs.append(mc_softmax_predictive(logits)[1].cpu().numpy())
    ood_probs = np.concatenate(ood_probs, axis=0)
    scores = np.concatenate([ood_scores(id_probs), ood_scores(ood_probs)])
    labels = np.concatenate([np.zeros(len(id_probs)), np.ones(len(ood_probs))])
    return float(roc_auc_score(labels, scores))
