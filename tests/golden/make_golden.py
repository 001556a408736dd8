#!/usr/bin/env python
"""Generates tests/golden/hotpath_v1.npz — committed input/output vectors for the hot path.

Why this script and not the reference itself: /root/reference is JAX/flax/matfree code and none of those
packages exist in this image, so the reference cannot be imported to emit vectors.  The vectors
are therefore produced by the float64 CPU oracle (oracle/lip_oracle.py), AFTER that oracle has been pinned
against the JAX-free golden numbers the reference's own tests hold (G1-G3 below, copied as literals with their
reference file:line) — tests/test_oracle_golden.py re-checks both on every CPU run, and the `-m gpu` tests compare
the CUDA path with the committed arrays without touching the oracle.

Run from the repo root:   python tests/golden/make_golden.py
"""
import math
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from oracle import lip_oracle as O   # noqa: E402
from oracle import models as OM      # noqa: E402

OUT = os.path.join(HERE, "hotpath_v1.npz")

# name: (kind, hidden, n_out, in_dim, M, N, logvar, model seed, data seed)
CASES = {
    "C1": ("regressor", [8, 8, 8, 8], 1, 1, 40, 240, 0.0, 100, 1000),     # toy sine regressor shape
    "C2": ("classifier", [16, 16], 2, 2, 32, 800, 0.0, 100, 1000),        # XOR classifier shape (C2)
    "RG": ("large", [37, 129, 20], 7, 45, 133, 1000, 0.0, 100, 1000),     # ragged LargeClassifier
}


def make_state(kind, hidden, n_out, in_dim, seed, logvar):
    if kind == "regressor":
        om = OM.OracleModel("regressor_mlp", (in_dim,), list(hidden), 1, "regressor")
    elif kind == "classifier":
        om = OM.OracleModel("classifier_mlp", (in_dim,), list(hidden), n_out, "classifier")
    else:
        om = OM.OracleModel("large_classifier", (in_dim,), list(hidden), n_out, "classifier")
    return OM.OracleState(om, om.init(seed), logvar=logvar)


def build():
    out = {}
    # ---- reference literals (JAX-free golden numbers of the reference's own tests) ----
    out["G1_X"] = np.array([[-1.0], [0.0], [1.1], [3.5]])                      # tests/fixtures.py:24
    out["G1_GGN_over_exp_neg_logvar"] = np.array([[4.0, 3.6], [3.6, 14.46]])   # tests/test_ggn.py:87-102 (bias, kernel order)
    out["G2_diag"] = np.arange(1, 101, dtype=np.float64) / 100                 # tests/test_sample.py:334-355
    out["G3_M1_trace"] = np.array(6.0)                                         # tests/fixtures.py:201
    out["G3_A"] = np.array([[1.0, 4, 50], [-30, 4.0, 16], [12, 6, 5.0]])       # tests/fixtures.py:205-209
    out["G3_M2_trace"] = np.array(3894.0)

    for name, (kind, hidden, n_out, in_dim, M, N, logvar, mseed, dseed) in CASES.items():
        st = make_state(kind, hidden, n_out, in_dim, mseed, logvar)
        mt = "regressor" if kind == "regressor" else "classifier"
        rng = np.random.default_rng(dseed)
        Z = rng.standard_normal((M, in_dim)).astype(np.float32)
        theta = st.flat()[0].astype(np.float32)
        D = theta.size
        prng = np.random.default_rng(dseed + 7)
        V = prng.choice([-1.0, 1.0], size=(4, D)).astype(np.float32)
        V[2:] = prng.standard_normal((2, D)).astype(np.float32)
        alpha = 0.37
        vp = O.compute_ggn_vp(st, Z, mt, full_set_size=N)
        Wf, WTf = O.compute_W_vps(st, Z, mt, full_set_size=N)
        wt = np.stack([np.asarray(WTf(v)) for v in V.astype(np.float64)])
        U = prng.standard_normal(wt.shape).astype(np.float32)
        out[f"{name}_theta"] = theta
        out[f"{name}_Z"] = Z
        out[f"{name}_V"] = V
        out[f"{name}_U"] = U
        out[f"{name}_meta"] = np.array([M, N, D, n_out], dtype=np.int64)
        out[f"{name}_alpha"] = np.array(alpha)
        out[f"{name}_logits"] = O.model_outputs(st, Z)
        out[f"{name}_ggn_vp"] = np.stack([vp(v) for v in V.astype(np.float64)])
        out[f"{name}_curvature_vp"] = out[f"{name}_ggn_vp"] + alpha * V.astype(np.float64)
        out[f"{name}_WT"] = wt
        out[f"{name}_W"] = np.stack([Wf(u) for u in U.astype(np.float64)])
        if name == "C2":
            # estimators / Krylov stage on the XOR-shaped classifier, identical probes
            eps = prng.choice([-1.0, 1.0], size=(64, D)).astype(np.float32)
            out["C2_eps"] = eps
            cvp = O.compute_curvature_approx(st, Z, mt, 0.9, full_set_size=N)
            out["C2_hutchinson"] = np.array(O.stochastic_trace_estimator_mvp(cvp, eps.astype(np.float64)))
            out["C2_hutchpp_v2"] = np.array(O.hutchpp_v2(cvp, eps.astype(np.float64), s1=48, s2=16))
            cvp17 = O.compute_curvature_approx(st, Z, mt, 1.7, full_set_size=N)
            out["C2_slq_lanczos_k25_clip1"] = np.array(O.slq_logdet_lanczos(cvp17, eps[:3].astype(np.float64), 25, clip_min=1.0))
            out["C2_slq_gkl_k25"] = np.array(O.slq_logdet_gkl(st, Z, mt, 1.7, eps[:3].astype(np.float64), 25))
            b = prng.standard_normal((3, D)).astype(np.float32)
            out["C2_cg_b"] = b
            sols = [O.cg(cvp, bb.astype(np.float64)) for bb in b]
            out["C2_cg_x"] = np.stack([s[0] for s in sols])
            out["C2_cg_iters"] = np.array([s[1] for s in sols], dtype=np.int64)
            Zs = Z[:12]
            Eps = prng.standard_normal((3, D)).astype(np.float32)
            Xnew = prng.standard_normal((9, in_dim)).astype(np.float32)
            out["C2_sample_eps"] = Eps
            out["C2_sample_Xnew"] = Xnew
            out["C2_sample"] = O.sample(st, Zs, D, 2.5, Eps.astype(np.float64), mt, full_set_size=N)
            out["C2_predict"] = O.predict_lla_scalable(st, Xnew, Zs, mt, 2.5, Eps.astype(np.float64), full_set_size=N)
            Wo, WTo = O.compute_W_vps(st, Z, mt)
            out["C2_gram"] = O.build_WTW(Wo, WTo, (M, n_out), M * n_out)
    # Lanczos-20 inverse square root of G2 (unclipped eigh, tests/test_sample.py:337)
    f = O.funm_lanczos_sym(O.dense_funm_sym_eigh(lambda x: 1.0 / np.sqrt(x), clip_min=None), O.tridiag_sym(20))
    out["G2_lanczos20_invsqrt_ones"] = f(lambda v: out["G2_diag"] * v, np.ones(100))
    return out


if __name__ == "__main__":
    data = build()
    np.savez_compressed(OUT, **data)
    print(f"wrote {OUT}: {len(data)} arrays, {os.path.getsize(OUT)} bytes")
