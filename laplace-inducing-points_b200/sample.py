"""Drop-in mirror of /root/reference/src/sample.py (matrix-free part) on the B200 path."""
from __future__ import annotations

import math

import torch

from ._runtime import dev_f32
from .ggn import _batched, build_WTW, compute_W_vps
from .matfree import _generator, decomp, dense_sym_operator, funm_lanczos_sym
from .matfree_monkeypatch import dense_funm_sym_eigh


def inv_matsqrt_vp(state, Z, D, alpha, model_type, full_set_size=None, key=None, num_proj_steps=1):
    """sample.py:55-145:  A^{-1/2} v for A = alpha I + beta W W^T (Higham et al. low-rank update):
         v -> W (W^T W)^{-1} (alpha I + beta W^T W)^{-1/2} W^T v + alpha^{-1/2} (v - W (W^T W)^{-1} W^T v)
    with the inverse square root from 2M Lanczos steps on the d x d Gram and eigenvalues clipped to >= 1
    (sample.py:113-115, matfree_monkeypatch.py:19).  `key` must be None (sample.py:150 forces the direct
    projection; the alternating-projection branch is dead code that returns NaN)."""
    if key is not None:
        raise NotImplementedError("alternating projections (sample.py:87-102) are dead code in the reference")
    Wfun, WTfun = compute_W_vps(state, Z, model_type, full_set_size=None)   # beta applied below (sample.py:63)
    bm = Wfun._lip_model
    dummy = WTfun(torch.zeros(D, device=bm.device))
    inner_shape, d = tuple(dummy.shape), dummy.numel()
    WTW = build_WTW(Wfun, WTfun, inner_shape, d, dtype=torch.float32, block=2)
    # jax.scipy.linalg.solve(WTW, .) (sample.py:81,135): LU of the (for classifiers singular) Gram, float64 here;
    # factorised once instead of once per call.
    LU, piv = torch.linalg.lu_factor(WTW.double())

    def solve(U):                       # U [B, d] -> [B, d]
        return torch.linalg.lu_solve(LU, piv, U.double().T).T.float()

    M = bm.M
    N = full_set_size or M
    beta = N / M
    invsqrt_fun = dense_funm_sym_eigh(lambda x: 1.0 / torch.sqrt(x))       # clipped (monkeypatched) version
    if 2 * M > d:
        raise ValueError(f"tridiag_sym(2*M={2 * M}) exceeds the Gram dimension d={d} (regressors: SURVEY §3.4)")
    invmatsqrt = funm_lanczos_sym(invsqrt_fun, decomp.tridiag_sym(2 * M))

    # sample.py:120-125: u -> alpha u + beta WTW u, the dense d x d mat-vec inside the Lanczos recurrence (LIP_LINOP_DENSE_SYM:
    # the whole 2M-step recurrence runs in one native call)
    inner_fun_flat = dense_sym_operator(WTW, alpha, beta)

    def vp(v):
        V = dev_f32(v)
        single = V.dim() == 1
        Vb = V.reshape(-1, D)
        u = WTfun(Vb).reshape(Vb.shape[0], d)                       # W^T v
        t = invmatsqrt(inner_fun_flat, u)                           # (alpha I + beta WTW)^{-1/2} W^T v
        x = solve(torch.cat([t, u], dim=0))                         # both solves share one call
        xt, xu = x[:Vb.shape[0]], x[Vb.shape[0]:]
        # outer_fun + alpha^{-1/2} nullproj:  W(xt) + a (v - W(xu)) = a v + W(xt - a xu)
        a = 1.0 / math.sqrt(alpha)
        comb = (xt - a * xu).reshape((Vb.shape[0],) + inner_shape)
        out = bm.w(comb, scale=Wfun._lip_scale, add=Vb, add_scale=a, batched=True)
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
