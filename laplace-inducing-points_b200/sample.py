"""Drop-in mirror of /root/reference/src/sample.py (matrix-free part) on the B200 path."""
from __future__ import annotations

import math

import torch

from ._runtime import dev_f32
from .ggn import _batched, build_WTW, compute_W_vps
from .matfree import DenseFunm, _generator, decomp, dense_sym_operator, funm_lanczos_sym


# Relative eigenvalue cut of the Gram's pseudo-inverse: directions of W^T W below PINV_TAU * lambda_max are treated as null(W).
# fp32 (3xTF32) Gram entries carry ~1e-6 relative noise, so nothing below that level can be told from an exact null direction.
PINV_TAU = 1e-6


def inv_matsqrt_vp(state, Z, D, alpha, model_type, full_set_size=None, key=None, num_proj_steps=1):
    """sample.py:55-145:  A^{-1/2} v for A = alpha I + beta W W^T (Higham et al. low-rank update):
         v -> W (W^T W)^{-1} f(alpha I + beta W^T W) W^T v + alpha^{-1/2} (v - W (W^T W)^{-1} W^T v),   f = clip(., 1)^{-1/2}
    with f(..) from 2M Lanczos steps on the d x d Gram (sample.py:113-115, matfree_monkeypatch.py:19).

    Re-arranged for the device, same mathematics:  A^{-1/2} v = alpha^{-1/2} v + W psi(W^T W) y,  y = W^T v,
         psi(lam) = (f(alpha + beta lam) - alpha^{-1/2}) / lam,
    so ONE Lanczos recurrence (lip_lanczos_tridiag on the dense Gram operator, LIP_LINOP_DENSE_SYM) + one tridiagonal
    eigen-decomposition (lip_tridiag_funm_p, LIP_FN_SAMPLER) deliver both terms; the reference's two LU solves with the
    rank-deficient Gram (sample.py:81,135 — each softmax L_i drops a direction, and the Jacobians' singular values decay below
    fp32 resolution) are the 1/lam of psi, applied as a pseudo-inverse with the cut PINV_TAU instead of dividing rounding noise by
    near-zero pivots.  `key` must be None (sample.py:150 forces the direct projection; the alternating-projection branch is dead
    code that returns NaN)."""
    if key is not None:
        raise NotImplementedError("alternating projections (sample.py:87-102) are dead code in the reference")
    Wfun, WTfun = compute_W_vps(state, Z, model_type, full_set_size=None)   # beta applied below (sample.py:63)
    bm = Wfun._lip_model
    inner_shape = (bm.M,) if model_type == "regressor" else (bm.M, bm.K)
    d = bm.M * bm.K
    WTW = build_WTW(Wfun, WTfun, inner_shape, d, dtype=torch.float32, block=2)
    M = bm.M
    N = full_set_size or M
    beta = N / M
    alpha = float(alpha)
    if 2 * M > d:
        raise ValueError(f"tridiag_sym(2*M={2 * M}) exceeds the Gram dimension d={d} (regressors: SURVEY §3.4)")
    psi = DenseFunm("sampler", clip_min=1.0, params=(alpha, beta, PINV_TAU))      # clipped as the monkeypatched dense_funm_sym_eigh
    output_space_term = funm_lanczos_sym(psi, decomp.tridiag_sym(2 * M))
    # sample.py:120-125: u -> alpha u + beta WTW u, the dense d x d mat-vec inside the Lanczos recurrence
    inner_fun_flat = dense_sym_operator(WTW, alpha, beta)

    def vp(v):
        V = dev_f32(v)
        single = V.dim() == 1
        Vb = V.reshape(-1, D)
        y = WTfun(Vb).reshape(Vb.shape[0], d)                       # W^T v
        z = output_space_term(inner_fun_flat, y)                    # psi(W^T W) W^T v
        out = bm.w(z.reshape((Vb.shape[0],) + inner_shape), scale=Wfun._lip_scale, add=Vb, add_scale=1.0 / math.sqrt(alpha),
                   batched=True)                                    # alpha^{-1/2} v + W z
        return out[0] if single else out

    return _batched(vp, bm, _lip_kind="INVSQRT")


def sample(state, Z, D, alpha, key, model_type, num_samples=1, full_set_size=None, num_proj_steps=10, *, eps=None):
    """sample.py:148-156: S zero-mean draws A^{-1/2} eps (the MAP is not added, :153-155); returns [S, D].
    The reference's sequential lax.map over samples is one batched call."""
    if eps is None:
        g = _generator(key)
        eps = torch.randn(num_samples, D, generator=g, device=g.device)
    inv_matsqrt_fun = inv_matsqrt_vp(state, Z, D, alpha, model_type, full_set_size=full_set_size, key=None,
                                     num_proj_steps=num_proj_steps)
    return inv_matsqrt_fun(dev_f32(eps).reshape(-1, D))
