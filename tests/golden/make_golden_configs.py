#!/usr/bin/env python
"""Generates tests/golden/configs_v2.npz — ground truth for the BASELINE configurations AT THEIR OWN SIZE
(VERDICT round 1, item 1): numbers the `-m gpu` parity tests compare the CUDA path with, without running the
(slow, CPU) oracle on the GPU box.

Two kinds of ground truth are stored, and the file says which is which:

  * `dense_*`  INDEPENDENT of the matfree / JAX-CG restatement: explicit per-point Jacobians (torch.func.jacrev through the
    float64 model), explicit W = [J_i^T L_i], dense eigendecomposition of A = alpha I + beta W W^T in float64.  Quadratic
    forms v^T log(A) v, v^T log(clip(A, 1)) v, logdet, A^{-1/2} v come from that decomposition only.
  * `oracle_*` the float64 oracle's own Lanczos / GKL / Hutch++ (oracle/lip_oracle.py) at sizes where a dense
    decomposition is out of reach (C3b: D = 1,494,154).

Inputs (weights, points, probes) are regenerated from the seeds below by the tests — numpy Generator streams are
stable — so only outputs are committed.

Run from the repo root (takes ~30 min on 8 cores):   python tests/golden/make_golden_configs.py [case ...]
"""
import math
import os
import sys
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from oracle import lip_oracle as O   # noqa: E402
from oracle import models as OM      # noqa: E402

OUT = os.path.join(HERE, "configs_v2.npz")


def rademacher(rng, shape):
    return (rng.integers(0, 2, size=shape, dtype=np.int8) * 2 - 1).astype(np.float64)


# ---------------------------------------------------------------------------------------------------------------------
# shared builders (the tests import these so that both sides see identical inputs)
# ---------------------------------------------------------------------------------------------------------------------
def c3a_inputs():
    """C3a: SimpleClassifier(32,3,2), D = 2274, M = 512 points ~ N(0,1)^2, N = 10000, alpha = 0.5,
    SLQ k = int(0.8 M) = 409, 4 Rademacher probes (main.py:186)."""
    om = OM.OracleModel("classifier_mlp", (2,), [32, 32, 32], 2, "classifier")
    ost = OM.OracleState(om, om.init(1003))
    rng = np.random.default_rng(2003)
    Z = rng.standard_normal((512, 2)).astype(np.float32)
    D = ost.flat()[0].size
    probes = rademacher(rng, (4, D))
    eps = rng.standard_normal((3, D)).astype(np.float32).astype(np.float64)
    return ost, Z, dict(alpha=0.5, N=10000, k=409), probes, eps


def c3b_inputs():
    """C3b headline (config/scale/mlp_mnist.yml): LargeClassifier 784-1024-512-256-128-10, M = 512, N = 60000, alpha = 1e-3."""
    om = OM.OracleModel("large_classifier", (28, 28, 1), [1024, 512, 256, 128], 10, "classifier")
    ost = OM.OracleState(om, om.init(1003))
    rng = np.random.default_rng(2004)
    Z = rng.random((512, 784)).astype(np.float32)
    D = ost.flat()[0].size
    slq_probes = rademacher(rng, (2, D))
    hpp_probes = rademacher(rng, (64, D))
    return ost, Z, dict(alpha=1e-3, N=60000, k=64, s1=48, s2=16), slq_probes, hpp_probes


def resnet_inputs():
    """C5 shape at M = 64 (d = 640): ResNet1M on CIFAR-shaped points, alpha = 5e-3, N = 49000 (resnet1-2_cifar10.yml)."""
    om = OM.ResNet1M(10, (32, 32, 3))
    ost = OM.OracleState(om, om.init(1005))
    rng = np.random.default_rng(2005)
    Z = rng.random((64, 32, 32, 3)).astype(np.float32)
    Xnew = rng.random((6, 32, 32, 3)).astype(np.float32)
    D = ost.flat()[0].size
    eps = rng.standard_normal((2, D)).astype(np.float32).astype(np.float64)
    return ost, Z, Xnew, dict(alpha=5e-3, N=49000), eps


def dense_factor(ost, Z):
    """Explicit W = [J_1^T L_1 ... J_M^T L_M] in R^{D x MK} (ggn.py:23-27,79-93 with full_set_size=None), float64."""
    J = O.jacobians(ost, Z)                                   # [M, K, D]
    p = O.softmax_np(O.model_outputs(ost, Z))
    sp = np.sqrt(p)
    cols = [J[i].T @ (np.diag(sp[i]) - np.outer(p[i], sp[i])) for i in range(J.shape[0])]
    return np.concatenate(cols, axis=1)


# ---------------------------------------------------------------------------------------------------------------------
def case_c3a(out):
    ost, Z, cfg, probes, eps = c3a_inputs()
    alpha, N, M = cfg["alpha"], cfg["N"], Z.shape[0]
    beta = N / M
    t0 = time.time()
    W = dense_factor(ost, Z)
    D = W.shape[0]
    lam, V = np.linalg.eigh(W @ W.T)                          # eigen-pairs of W W^T (float64)
    lam = np.clip(lam, 0.0, None)
    c = probes @ V                                            # [4, D] coordinates of the probes
    c2 = c * c
    # production logdet (train_inducing.py:156-171): A1 = alpha I + W W^T (no beta)
    a1 = alpha + lam
    out["c3a_dense_logdet_gkl"] = np.log(a1).sum()
    out["c3a_dense_quad_log_gkl"] = c2 @ np.log(a1)
    # Lanczos form on curvature_vp (train_inducing.py:152-153): A2 = alpha I + beta W W^T, patched integrand clips eigenvalues at 1
    a2 = alpha + beta * lam
    out["c3a_dense_quad_logclip_lanczos"] = c2 @ np.log(np.clip(a2, 1.0, None))
    out["c3a_dense_quad_log_lanczos"] = c2 @ np.log(a2)
    # A2^{-1/2} eps.  `true`: the matrix function itself.  `clip`: the exact-arithmetic value of the reference's formula
    # (sample.py:117-143), whose Lanczos funm clips the eigenvalues of alpha I + beta W^T W at 1 (matfree_monkeypatch.py:19) on
    # range(W) and uses alpha^{-1/2} on its complement.
    rng_dirs = lam > 1e-9 * lam.max()
    g_true = 1.0 / np.sqrt(a2)
    g_clip = np.where(rng_dirs, 1.0 / np.sqrt(np.clip(a2, 1.0, None)), 1.0 / math.sqrt(alpha))
    ce = eps @ V
    out["c3a_dense_invsqrt_true"] = (ce * g_true) @ V.T
    out["c3a_dense_invsqrt_clip"] = (ce * g_clip) @ V.T
    out["c3a_rank"] = np.array(int(rng_dirs.sum()))
    # the same with the pseudo-inverse cut an fp32 implementation has to use (directions below 1e-6 lambda_max count as null(W));
    # for alpha < 1 the clipped formula is DISCONTINUOUS at that cut (1 on range(W), alpha^{-1/2} on the complement), so which side
    # a marginal direction falls on changes the result at the 1e-2 level — no fp32 method can be pinned tighter than that there
    rng6 = lam > 1e-6 * lam.max()
    g_clip6 = np.where(rng6, 1.0 / np.sqrt(np.clip(a2, 1.0, None)), 1.0 / math.sqrt(alpha))
    out["c3a_dense_invsqrt_clip_tau1e-6"] = (ce * g_clip6) @ V.T
    out["c3a_rank_tau1e-6"] = np.array(int(rng6.sum()))
    # alpha = 1: the clip is inactive (alpha + beta lam >= 1), the reference's formula IS A^{-1/2} and is continuous in lam
    a3 = 1.0 + beta * lam
    out["c3a_dense_invsqrt_alpha1"] = (ce / np.sqrt(a3)) @ V.T
    print(f"[c3a] D={D} rank={int(rng_dirs.sum())} logdet={out['c3a_dense_logdet_gkl']:.6f} ({time.time() - t0:.1f} s)", flush=True)


def case_c3b_slq(out):
    ost, Z, cfg, slq_probes, _ = c3b_inputs()
    t0 = time.time()
    vals = [O.slq_logdet_gkl(ost, Z, "classifier", cfg["alpha"], slq_probes[i:i + 1], cfg["k"]) for i in range(slq_probes.shape[0])]
    out["c3b_oracle_gkl_quad_k64"] = np.array(vals)
    print(f"[c3b] GKL k=64: {vals} ({time.time() - t0:.1f} s)", flush=True)
    t0 = time.time()
    cvp = O.compute_curvature_approx(ost, Z, "classifier", cfg["alpha"], full_set_size=cfg["N"])
    vals = [O.slq_logdet_lanczos(cvp, slq_probes[i:i + 1], cfg["k"], clip_min=1.0) for i in range(slq_probes.shape[0])]
    out["c3b_oracle_lanczos_quad_k64"] = np.array(vals)
    print(f"[c3b] Lanczos k=64: {vals} ({time.time() - t0:.1f} s)", flush=True)


def case_c3b_hpp(out):
    ost, Z, cfg, _, hpp_probes = c3b_inputs()
    t0 = time.time()
    cvp = O.compute_curvature_approx(ost, Z, "classifier", cfg["alpha"], full_set_size=cfg["N"])
    out["c3b_oracle_hutchpp_v2"] = np.array(O.hutchpp_v2(cvp, hpp_probes, s1=cfg["s1"], s2=cfg["s2"]))
    out["c3b_oracle_hutchinson"] = np.array(O.stochastic_trace_estimator_mvp(cvp, hpp_probes[:16]))
    print(f"[c3b] hutchpp_v2={out['c3b_oracle_hutchpp_v2']} hutchinson16={out['c3b_oracle_hutchinson']} ({time.time() - t0:.1f} s)",
          flush=True)


def case_resnet(out):
    ost, Z, Xnew, cfg, eps = resnet_inputs()
    D = ost.flat()[0].size
    t0 = time.time()
    w = O.sample(ost, Z, D, cfg["alpha"], eps, "classifier", full_set_size=cfg["N"])
    out["resnet_oracle_sample_strided"] = w[:, ::257].copy()     # every 257th entry (the full [2, D] block would be 17 MB)
    out["resnet_oracle_sample_norm"] = np.linalg.norm(w, axis=1)
    print(f"[resnet] sample ({time.time() - t0:.1f} s)", flush=True)
    f = O._model_fn(ost, Xnew)
    import torch
    th = O._t(ost.flat()[0])
    fmu = f(th).detach().numpy()
    dys = np.stack([torch.func.jvp(f, (th,), (O._t(ws),))[1].detach().numpy() for ws in w])
    out["resnet_oracle_predict"] = fmu[None] + dys
    print(f"[resnet] predict ({time.time() - t0:.1f} s)", flush=True)


CASES = {"c3a": case_c3a, "c3b_slq": case_c3b_slq, "c3b_hpp": case_c3b_hpp, "resnet": case_resnet}


def main():
    want = sys.argv[1:] or list(CASES)
    for name in want:
        new = {}
        CASES[name](new)
        out = dict(np.load(OUT)) if os.path.exists(OUT) else {}       # merge with whatever is on disk NOW (cases may run concurrently)
        out.update(new)
        np.savez_compressed(OUT, **out)
    print("wrote", OUT, sorted(out))


if __name__ == "__main__":
    main()
