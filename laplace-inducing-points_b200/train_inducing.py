"""Forward value of the reference's scalable KL objective on the B200 path (/root/reference/src/train_inducing.py:87-173).

This is the production CALLER of the hot path: every heavy step is one of the package's operators —
curvature_vp over the data minibatch, W_z / W_z^T, the dense Gram, Hutch++ v2, the GKL logdet.  Only the forward value is
provided: differentiating it w.r.t. Z (train_inducing.py:196, row f1) is not built, so this is an evaluation / monitoring
entry point, not a training step.
"""
from __future__ import annotations

import math

import torch

from . import matfree
from ._runtime import dev_f32
from .ggn import build_WTW, compute_W_vps
from .lla import compute_curvature_approx
from .stochtrace import hutchpp_v2
from .utils import flatten_nn_params


def alternative_objective_scalable(Z, X, state, alpha, model_type, key, full_set_size=None, st_samples=256, slq_samples=2,
                                   slq_num_matvecs=None, *, probes=None):
    """KL[q(theta|Z) || q(theta|data)] up to constants = tr(S_X S_Z^{-1}) + logdet(S_Z)   (train_inducing.py:87-173).

    Same arguments as the reference; `key` seeds the Rademacher probes (int / torch.Generator) unless `probes`
    [st_samples, D] is given (JAX's threefry stream is not reproducible here; probes are inputs)."""
    N = full_set_size
    Zt, Xt = dev_f32(Z), dev_f32(X)
    M = int(Zt.shape[0])
    beta = N / M
    alpha = float(alpha)
    alpha_inv, beta_inv = 1.0 / alpha, 1.0 / beta
    flat, _ = flatten_nn_params(state.params)          # D excludes logvar, as :104-106
    D = int(flat.numel())

    S_vp = compute_curvature_approx(state, Xt, model_type, alpha, full_set_size=N)               # :108-110
    # (the reference also builds Sz_vp over Z, :111-113, and never uses it)
    Wz, WzT = compute_W_vps(state, Zt, model_type, full_set_size=None)                           # :114-116
    bm = Wz._lip_model
    inner_shape = (M,) if model_type == "regressor" else (M, bm.K)
    d_z = M * bm.K
    WzTWz = build_WTW(Wz, WzT, inner_shape, d_z, dtype=torch.float32, block=1)                   # :126
    # Woodbury: S_Z^{-1} v = v/alpha - alpha^-2 Wz (beta^-1 I + alpha^-1 WzTWz)^-1 WzT v   (:127-132); the d_z x d_z system is
    # factorised once (float64 LU, library call) instead of once per matvec
    Kmat = beta_inv * torch.eye(d_z, device=WzTWz.device, dtype=torch.float64) + alpha_inv * WzTWz.double()
    LU, piv = torch.linalg.lu_factor(Kmat)

    @matfree.batched
    def Sz_inv(V):
        V = dev_f32(V).reshape(-1, D)
        u = WzT(V).reshape(V.shape[0], d_z)
        x = torch.linalg.lu_solve(LU, piv, u.double().T).T.float()
        return bm.w(x.reshape((V.shape[0],) + inner_shape), scale=Wz._lip_scale, add=V, add_scale=-alpha,
                    batched=True).mul_(-alpha_inv ** 2)      # -(1/alpha^2) (Wz x - alpha v) = v/alpha - Wz x / alpha^2

    @matfree.batched
    def composite_vp(V):                                                                         # :134-135
        return S_vp(Sz_inv(V))

    if probes is None:                                                                           # :138-142
        probes = matfree.sampler_rademacher(torch.ones(D), num=st_samples)(key)
    probes = dev_f32(probes)
    st_samples = int(probes.shape[0])
    trace_term = hutchpp_v2(composite_vp, lambda _: probes, s1=st_samples - 16, s2=16)           # :144-145

    k = slq_num_matvecs if slq_num_matvecs is not None else int(M * 0.8)                         # :148
    sqrt_alpha = math.sqrt(alpha)

    @matfree.batched
    def bidiag_target(V):                                                                        # :166-169
        V = V.reshape(-1, D)
        return torch.cat([sqrt_alpha * V, WzT(V).reshape(V.shape[0], d_z)], dim=1)

    @matfree.batched
    def bidiag_target_T(U):                                   # jax.vjp of bidiag_target inside matfree.decomp.bidiag
        U = U.reshape(-1, D + d_z)
        return Wz(U[:, D:].reshape((-1,) + inner_shape)).add_(U[:, :D], alpha=sqrt_alpha)

    problem = matfree.funm.integrand_funm_product_logdet(matfree.decomp.bidiag(k))               # :156-157
    logdet_term = problem(bidiag_target, probes[:slq_samples], bidiag_target_T).mean()           # :159-163
    return logdet_term + trace_term
