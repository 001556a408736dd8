#!/usr/bin/env python
"""Generates tests/golden/zgrad_v1.npz — committed vectors for SURVEY §8 rows f1 / f3: gradients of the hot-path operators
and of the deterministic objectives with respect to the inducing points Z, and the alpha evidence.

Produced by the float64 CPU oracle (torch.func.grad over a literal restatement of src/ggn.py / src/train_inducing.py:26-84,
175-192 / src/train_alpha.py:13-44); the reference itself (JAX) cannot run in this image (see make_golden.py).
tests/test_oracle_zgrad.py checks that oracle against central finite differences on every CPU run.

Run from the repo root:   python tests/golden/make_golden_zgrad.py
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, HERE)

from oracle import lip_oracle as O   # noqa: E402
from make_golden import make_state   # noqa: E402

OUT = os.path.join(HERE, "zgrad_v1.npz")

# name: (kind, hidden, n_out, in_dim, M, n_x, N, alpha, logvar, model seed, data seed)
CASES = {
    "ZC": ("classifier", [16, 16], 3, 2, 12, 20, 800, 0.05, 0.0, 200, 2000),
    "ZR": ("regressor", [8, 8], 1, 1, 10, 16, 240, 0.5, 0.3, 201, 2001),
}


def build():
    out = {}
    for name, (kind, hidden, n_out, in_dim, M, nx, N, alpha, logvar, mseed, dseed) in CASES.items():
        st = make_state(kind, hidden, n_out, in_dim, mseed, logvar)
        mt = "regressor" if kind == "regressor" else "classifier"
        rng = np.random.default_rng(dseed)
        Z = rng.standard_normal((M, in_dim)).astype(np.float32)
        X = rng.standard_normal((nx, in_dim)).astype(np.float32)
        theta = st.flat()[0].astype(np.float32)
        D = theta.size
        U = rng.standard_normal((3, D)).astype(np.float32)
        V = rng.standard_normal((3, D)).astype(np.float32)
        Y = rng.standard_normal((3, M, n_out)).astype(np.float32)
        out[f"{name}_theta"], out[f"{name}_Z"], out[f"{name}_X"] = theta, Z, X
        out[f"{name}_U"], out[f"{name}_V"], out[f"{name}_Y"] = U, V, Y
        out[f"{name}_meta"] = np.array([M, nx, N, D, n_out], dtype=np.int64)
        out[f"{name}_alpha"] = np.array(alpha)
        out[f"{name}_ggn_zgrad"] = O.ggn_vp_zgrad(st, Z, mt, U, V, full_set_size=N, per_probe=True)
        Wz, WTz = O.W_vps_zgrad(st, Z, mt, full_set_size=N)
        out[f"{name}_W_zgrad"] = Wz(U, Y)
        out[f"{name}_WT_zgrad"] = WTz(Y, V)
        v, g = O.variational_grad_scalable_exact(Z, X, st, alpha, mt, full_set_size=N)
        out[f"{name}_exact_value"], out[f"{name}_exact_zgrad"] = np.array(v), g
        v, g = O.variational_grad_dense(Z, X, st, alpha, mt, full_set_size=N)
        out[f"{name}_dense_value"], out[f"{name}_dense_zgrad"] = np.array(v), g
        out[f"{name}_lml"] = np.array(O.log_marginal_likelihood(alpha, X, st, mt, full_set_size=N))
    return out


if __name__ == "__main__":
    data = build()
    np.savez_compressed(OUT, **data)
    print(f"wrote {OUT}: {len(data)} arrays, {os.path.getsize(OUT)} bytes")
