"""The reference's inducing-point objectives and their gradients with respect to Z on the B200 path
(/root/reference/src/train_inducing.py).

This is the production CALLER of the hot path: every heavy step is one of the package's operators —
curvature_vp over the data minibatch, W_z / W_z^T, the dense / cross Gram, Hutch++ v2, the GKL logdet — and the Z-gradients
are assembled from lip_zgrad (the VJP-with-respect-to-Z rules of those operators; csrc/lip_zgrad.cu).

  alternative_objective_scalable        :87-173   forward value (Hutch++ v2 + GKL logdet)
  alternative_objective_scalable_exact  :26-84    value;  variational_grad_scalable_exact -> (value, dZ)   [deterministic]
  alternative_objective_dense           :175-192  value;  variational_grad_dense          -> (value, dZ)   [deterministic]
  variational_grad_scalable             :195      (stochastic loss value, EXACT dZ of the quantity it estimates; see its docstring:
                                                  NOT the reference's autodiff-through-the-estimators)
  optimize_step                         :198-232  one optimiser step on Z
"""
from __future__ import annotations

import math

import torch

from . import matfree
from ._runtime import dev_f32
from .ggn import build_WTW, build_WTWz, compute_W_vps
from .lla import compute_curvature_approx
from .stochtrace import hutchpp_v2
from .utils import flatten_nn_params


def alternative_objective_scalable(Z, X, state, alpha, model_type, key, full_set_size=None, st_samples=256, slq_samples=2,
                                   slq_num_matvecs=None, *, probes=None, _parts=None):
    """KL[q(theta|Z) || q(theta|data)] up to constants = tr(S_X S_Z^{-1}) + logdet(S_Z)   (train_inducing.py:87-173).

    Same arguments as the reference; `key` seeds the Rademacher probes (int / torch.Generator) unless `probes`
    [st_samples, D] is given (JAX's threefry stream is not reproducible here; probes are inputs)."""
    N = full_set_size
    Zt, Xt = dev_f32(Z), dev_f32(X)
    M = int(Zt.shape[0])
    alpha = float(alpha)
    flat, _ = flatten_nn_params(state.params)          # D excludes logvar, as :104-106
    D = int(flat.numel())

    # S_vp over the minibatch (:108-110), W_z / W_z^T (:114-116), the dense Gram (:126) and the Woodbury inverse (:127-132); the d_z x d_z
    # system is factorised once (float64 LU, library call) instead of once per matvec.  (The reference also builds Sz_vp over Z,
    # :111-113, and never uses it.)
    parts = _parts if _parts is not None else _scalable_parts(Zt, Xt, state, alpha, model_type, N)
    S_vp, Wz, WzT, Sz_inv, inner_shape, d_z = (parts[k] for k in ("S_vp", "Wz", "WzT", "Sz_inv", "inner_shape", "d_z"))

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

    # bidiag_target v -> [sqrt(alpha) v ; W_z^T v] (:166-169) and its transpose (jax.vjp inside matfree.decomp.bidiag): the
    # closure pair carries the bound model, so the whole Golub-Kahan recurrence runs natively (LIP_LINOP_GKL)
    bidiag_target = matfree.gkl_target(WzT, Wz, alpha)
    bidiag_target_T = bidiag_target._lip_transpose

    problem = matfree.funm.integrand_funm_product_logdet(matfree.decomp.bidiag(k))               # :156-157
    logdet_term = problem(bidiag_target, probes[:slq_samples], bidiag_target_T).mean()           # :159-163
    return logdet_term + trace_term


# ============================================================================================================
# deterministic objectives with gradients (rows f1 / f3)
# ============================================================================================================
def _f64(x):
    return x.to(torch.float64)


_WS_BUDGET = 24 << 30        # bytes of operator workspace one block of probes may take (conv programs: [B, M, H, W, C] slots)


def _probe_block(D: int, want: int, models=(), zgrad_mode=None) -> int:
    """how many [D]-vectors to push through the operators at once: bounded probe block (1 GiB of floats) AND bounded operator
    workspace on every model in `models` (lip_workspace_bytes, and lip_zgrad_workspace_bytes when zgrad_mode is given) — residual
    conv programs keep several [B, M, H, W, C] tangent / cotangent slots, so 256 probes at M = 100 CIFAR images would ask for 60 GB."""
    from . import _cabi as cabi
    b = max(1, min(int(want), (1 << 28) // max(D, 1)))
    L = cabi.lib()

    def need(bm, nb):
        n = int(L.lip_workspace_bytes(bm._h, nb))
        if zgrad_mode is not None:
            n = max(n, int(L.lip_zgrad_workspace_bytes(bm._h, zgrad_mode, nb)))
        return n

    for bm in models:
        while b > 1 and need(bm, b) > _WS_BUDGET:
            b = max(1, b // 2)
    return b


_PSD_MAX = 4096
_ACCURATE_MAX_ELEMS = 12 * (1 << 30)     # floats of the materialised factor W_z (48 GB of the 180 GB HBM)


def _grams_from_factors(Wz, W, inner_shape_z, inner_shape_x, d_z, d, D):
    """G = W_z^T W_z and C = W^T W_z as float64 products of the MATERIALISED fp32 factors (rows W e_k from lip_w_apply).
    lip_gram_wtw returns fl32(W^T (W e_k)): its rounding error (~1e-6 |G|, either sign) lands directly on G's null / small eigenvalues.
    The float64 Gram of the rounded factor W~ is the exact Gram of a nearby matrix instead: positive semi-definite by construction, and
    its null directions are perturbed only to second order (|W~ n|^2 ~ 1e-12 |W|^2).  Measured (tools/grad_conditioning.py): 7x smaller
    gradient error at benign beta/alpha (3.8e-5 -> 5.3e-6), no gain at beta/alpha ~ 1e6, where the fp32 operators downstream of the
    Grams dominate - so it is an option (accurate_grams=True), not the default.
    The products are D-chunked float64 GEMMs (library calls, like the d_z x d_z factorisations next to them)."""
    dev = torch.device("cuda", torch.cuda.current_device())
    blk = _probe_block(D, 256)
    eye_z = torch.eye(d_z, device=dev, dtype=torch.float32)
    Wz_rows = torch.empty(d_z, D, device=dev, dtype=torch.float32)
    for k0 in range(0, d_z, blk):
        E = eye_z[k0:min(d_z, k0 + blk)]
        Wz_rows[k0:k0 + E.shape[0]] = Wz(E.reshape((-1,) + tuple(inner_shape_z))).reshape(E.shape[0], D)
    chunk = 1 << 18
    G = torch.zeros(d_z, d_z, device=dev, dtype=torch.float64)
    for c0 in range(0, D, chunk):
        A = Wz_rows[:, c0:c0 + chunk].double()
        G += A @ A.T
    C = torch.zeros(d, d_z, device=dev, dtype=torch.float64)
    eye_x = torch.eye(d, device=dev, dtype=torch.float32)
    for k0 in range(0, d, blk):
        E = eye_x[k0:min(d, k0 + blk)]
        Wx_rows = W(E.reshape((-1,) + tuple(inner_shape_x))).reshape(E.shape[0], D)
        for c0 in range(0, D, chunk):
            C[k0:k0 + E.shape[0]] += Wx_rows[:, c0:c0 + chunk].double() @ Wz_rows[:, c0:c0 + chunk].double().T
    return G, C


def _psd_part(G):
    """G = W_z^T W_z is positive semi-definite with one exact null direction per softmax point, but arrives with fp32 rounding noise
    (~1e-6 |G|) of either sign.  (I/beta + G/alpha)^-1 amplifies a negative noise eigenvalue without bound once it approaches
    -alpha/beta (the scale configs: alpha/beta ~ 1e-6), so the float64 small-matrix algebra works on the PSD part of the symmetrised
    Gram (eigenvalues clamped at 0: the inverse is then bounded by beta in every direction).  d_z <= 4096; larger Grams are used as is."""
    G = 0.5 * (G + G.T)
    if G.shape[0] > _PSD_MAX:
        return G
    lam, V = torch.linalg.eigh(G)
    return (V * lam.clamp_min(0.0)) @ V.T


def _exact_parts(Z, X, state, alpha, model_type, full_set_size, zside=None, accurate=False):
    """zside: optional (Wz, WzT, WzTWz) already built for the same state / Z (the scalable objective's parts)"""
    N = full_set_size
    Zt, Xt = dev_f32(Z), dev_f32(X)
    M, Kx = int(Zt.shape[0]), int(Xt.shape[0])
    flat, _ = flatten_nn_params(state.params)
    D = int(flat.numel())                                                                  # :37-39 (logvar is not in D)
    if zside is not None:
        Wz, WzT, WzTWz = zside
    else:
        Wz, WzT = compute_W_vps(state, Zt, model_type, full_set_size=None)                 # :48-50
        WzTWz = None
    W, WT = compute_W_vps(state, Xt, model_type, full_set_size=None)                       # :51-53
    bz, bx = Wz._lip_model, W._lip_model
    inner_shape = (M,) if model_type == "regressor" else (M, bz.K)
    d_z, d = M * bz.K, Kx * bx.K
    if accurate and (d_z + 256) * D <= _ACCURATE_MAX_ELEMS:
        WzTWz, WTWz = _grams_from_factors(Wz, W, inner_shape, (Kx,) if model_type == "regressor" else (Kx, bx.K), d_z, d, D)
    else:
        if WzTWz is None:
            WzTWz = build_WTW(Wz, WzT, inner_shape, d_z, dtype=torch.float32, block=1)     # :60
        WzTWz = _f64(WzTWz)
        WTWz = _f64(build_WTWz(WT, Wz, inner_shape, d=d, dtype=torch.float32, block=1))    # :67
    WzTWz = _psd_part(WzTWz)
    return dict(N=N, M=M, Kx=Kx, D=D, Wz=Wz, W=W, bz=bz, bx=bx, inner_shape=inner_shape, d_z=d_z, d=d, G=WzTWz, C=WTWz,
                beta=N / M, gamma=N / Kx, alpha=float(alpha))


def _exact_value(p):
    alpha, beta, gamma, G, C, d_z = p["alpha"], p["beta"], p["gamma"], p["G"], p["C"], p["d_z"]
    eye = torch.eye(d_z, device=G.device, dtype=torch.float64)
    logdet_term = torch.linalg.slogdet(eye + beta / alpha * G)[1] + p["D"] * math.log(alpha)    # :62-63
    Mm = eye / beta + G / alpha                                                            # :69
    # :70-72 factorise Mm by Cholesky.  G is the Gram of a rank-deficient factor (each softmax L_i drops one direction), so Mm's
    # smallest eigenvalues are exactly 1/beta and the fp32 rounding noise of G, amplified by 1/alpha, can push them below zero at the
    # scale configs (jnp.linalg.cholesky would return NaN there); an LU factorisation solves the same systems without that failure.
    L = torch.linalg.lu_factor(Mm)
    S1 = torch.linalg.lu_solve(*L, G)                                                      # :71
    S2 = torch.linalg.lu_solve(*L, C.T.contiguous())                                       # :72
    trace1 = torch.trace(S1)                                                               # :74
    trace2 = (C * S2.T).sum()                                                              # :75
    return logdet_term - trace1 / alpha - gamma / alpha ** 2 * trace2, L                   # :76-78


def alternative_objective_scalable_exact(Z, X, state, alpha, model_type, key=None, full_set_size=None, st_samples=256,
                                         slq_samples=2, slq_num_matvecs=None):
    """train_inducing.py:26-84: the KL objective through the dense Grams W_z^T W_z and W^T W_z (identity of
    src/Untitled-1.md:1-2), exact traces and slogdet.  The Grams come from lip_gram_wtw / lip_gram_cross; the d_z x d_z
    Cholesky / slogdet are library calls in float64, as the reference hands them to jnp.linalg."""
    return _exact_value(_exact_parts(Z, X, state, alpha, model_type, full_set_size))[0].float()


def variational_grad_scalable_exact(Z, X, state, alpha, model_type, key=None, full_set_size=None, *, accurate_grams=False, **_):
    """jax.value_and_grad of alternative_objective_scalable_exact with respect to Z -> (loss, dZ).

    With Mm = I/beta + G/alpha (G = W_z^T W_z, C = W^T W_z):  dL/dG = Mm^-1 G Mm^-1 / alpha^2 + gamma Mm^-1 C^T C Mm^-1 / alpha^3,
    dL/dC = -2 gamma C Mm^-1 / alpha^2, and  dL/dZ = sum_k d/dZ < 2 W_z (dL/dG)[:, k] + W (dL/dC)[:, k],  W_z e_k >  — one
    lip_zgrad(W mode) call per block of one-hot columns."""
    p = _exact_parts(Z, X, state, alpha, model_type, full_set_size, accurate=accurate_grams)
    value, dZ = _exact_value_and_zgrad(p)
    return value.float(), dZ.reshape(dev_f32(Z).shape)


def _exact_value_and_zgrad(p):
    value, L = _exact_value(p)
    alpha, gamma, G, C, d_z, D = p["alpha"], p["gamma"], p["G"], p["C"], p["d_z"], p["D"]
    eye = torch.eye(d_z, device=G.device, dtype=torch.float64)
    Minv = torch.linalg.lu_solve(*L, eye)
    Minv = 0.5 * (Minv + Minv.T)
    Gbar = (Minv @ G @ Minv) / alpha ** 2 + gamma / alpha ** 3 * (Minv @ (C.T @ C) @ Minv)
    Cbar = -2.0 * gamma / alpha ** 2 * (C @ Minv)
    Wz, W, bz = p["Wz"], p["W"], p["bz"]
    from . import _cabi as cabi
    blk = _probe_block(D, 256, models=(bz, W._lip_model), zgrad_mode=cabi.ZGRAD_W)
    dZ = torch.zeros(p["M"], bz.Z.shape[1], device=G.device, dtype=torch.float32)
    onehot = torch.eye(d_z, device=G.device, dtype=torch.float32)
    for k0 in range(0, d_z, blk):
        k1 = min(d_z, k0 + blk)
        nb = k1 - k0
        ub = W._lip_model.w(Cbar[:, k0:k1].T.float().contiguous(), scale=W._lip_scale, batched=True)         # W Cbar[:, k]
        ub = bz.w((2.0 * Gbar[:, k0:k1]).T.float().contiguous(), scale=Wz._lip_scale, add=ub, add_scale=1.0, batched=True)
        dZ += Wz.zgrad(ub, onehot[k0:k1].reshape(nb, d_z))
    return value, dZ


def _dense_parts(Z, X, state, alpha, model_type, full_set_size):
    from .ggn import compute_ggn_dense
    Zt, Xt = dev_f32(Z), dev_f32(X)
    S, flat, _ = compute_ggn_dense(state, Xt, model_type, full_set_size)                   # lla.py compute_curvature_approx_dense
    S_z, _, _ = compute_ggn_dense(state, Zt, model_type, full_set_size)
    D = int(flat.numel())
    eye = torch.eye(D, device=S.device, dtype=torch.float64)
    S = _f64(S) + alpha * eye
    S_z = _f64(S_z) + alpha * eye
    S_z_inv = torch.linalg.inv(S_z)                                                        # :184
    return S, S_z_inv, D


def alternative_objective_dense(Z, X, state, alpha, model_type, key=None, full_set_size=None):
    """train_inducing.py:175-192: tr(S S_z^-1) - logdet(S_z^-1) with the D x D matrices materialised (toy sizes)."""
    S, S_z_inv, _ = _dense_parts(Z, X, state, float(alpha), model_type, full_set_size)
    return (torch.trace(S @ S_z_inv) - torch.linalg.slogdet(S_z_inv)[1]).float()


def variational_grad_dense(Z, X, state, alpha, model_type, key=None, full_set_size=None, **_):
    """train_inducing.py:194: jax.value_and_grad(alternative_objective_dense) -> (loss, dZ).
    dL/dS_z = S_z^-1 - S_z^-1 S S_z^-1;  dL/dZ = sum_k d/dZ < (dL/dS_z)[:, k], GGN(Z) e_k >  (lip_zgrad, GGN mode)."""
    from .ggn import compute_ggn_vp
    alpha = float(alpha)
    S, S_z_inv, D = _dense_parts(Z, X, state, alpha, model_type, full_set_size)
    value = torch.trace(S @ S_z_inv) - torch.linalg.slogdet(S_z_inv)[1]
    Sbar = S_z_inv - S_z_inv @ S @ S_z_inv
    Sbar = 0.5 * (Sbar + Sbar.T)
    Zt = dev_f32(Z)
    vp = compute_ggn_vp(state, Zt, model_type, full_set_size)
    eye = torch.eye(D, device=Zt.device, dtype=torch.float32)
    from . import _cabi as cabi
    blk = _probe_block(D, 512, models=(vp._lip_model,), zgrad_mode=cabi.ZGRAD_GGN)
    dZ = torch.zeros(Zt.shape[0], vp._lip_model.Z.shape[1], device=Zt.device, dtype=torch.float32)
    for k0 in range(0, D, blk):
        k1 = min(D, k0 + blk)
        dZ += vp.zgrad(Sbar[:, k0:k1].T.float().contiguous(), eye[k0:k1])
    return value.float(), dZ.reshape(Zt.shape)


def variational_grad_scalable(Z, X, state, alpha, model_type, key, full_set_size=None, st_samples=256, slq_samples=2,
                              slq_num_matvecs=None, *, probes=None, gradient="exact"):
    """(loss, dZ) for the scalable objective (train_inducing.py:195).

    The loss is alternative_objective_scalable (Hutch++ v2 + GKL logdet on the given probes).  The reference differentiates
    THROUGH those estimators (QR, Golub-Kahan recurrences, SVD) and so gets a noisy gradient of a noisy loss; here dZ is the gradient
    of the quantity they estimate, tr(S_X S_Z^-1) + logdet S_Z:

      gradient="exact" (default): the Gram-space formulas of variational_grad_scalable_exact (the two objectives differ by a Z-independent
        constant) on the Grams already built for the Woodbury inverse: float64 small-matrix algebra on the PSD part of the fp32 Grams,
        d_z one-hot columns through lip_zgrad(W mode).  Deterministic, no probe variance.
      gradient="woodbury": S_Z depends on Z only through the D x d_z factor W_z, so dL = tr(A dS_Z), A = S_Z^-1 - S_Z^-1 S_X S_Z^-1, is
        sum_k d/dZ < 2 beta A W_z e_k, W_z e_k > with A pushed through the fp32 Woodbury inverse (no cross-Gram needed; ~50x more rounding error).
      gradient="hutchinson": mean_b d/dZ < S_Z^-1 eps_b , GGN(Z) (eps_b - S_Z^-1 S_X eps_b) > on the loss's probes (ONE lip_zgrad call,
        probes sharded over ranks).  Unbiased but, for D ~ 1e6 and tens of probes, dominated by its variance.
      All three lose accuracy as beta / alpha grows (profiles/r01_descent_check.txt, tools/grad_conditioning.py).

    Parity with the reference is claimed for the deterministic forms (oracle: float64 autograd of train_inducing.py:26-84,175-192)."""
    from .ggn import compute_ggn_vp
    N = full_set_size
    Zt, Xt = dev_f32(Z), dev_f32(X)
    alpha = float(alpha)
    flat, _ = flatten_nn_params(state.params)
    D = int(flat.numel())
    if probes is None:
        probes = matfree.sampler_rademacher(torch.ones(D), num=st_samples)(key)
    probes = dev_f32(probes)
    parts = _scalable_parts(Zt, Xt, state, alpha, model_type, N)
    loss = alternative_objective_scalable(Zt, Xt, state, alpha, model_type, key, full_set_size=N, st_samples=st_samples,
                                          slq_samples=slq_samples, slq_num_matvecs=slq_num_matvecs, probes=probes, _parts=parts)
    if gradient == "exact":
        # Gram-space form (float64 small-matrix algebra on the PSD part of the fp32 Grams): 50x less rounding error than pushing A
        # through fp32 Woodbury solves (tools/grad_conditioning.py: 2.7e-3 vs 1.0e-1 relative at alpha = 1e-3, beta = 500 on a toy model)
        p = _exact_parts(Zt, Xt, state, alpha, model_type, N, zside=(parts["Wz"], parts["WzT"], parts["WzTWz"]))
        _, dZ = _exact_value_and_zgrad(p)
        return loss, dZ.reshape(Zt.shape)
    if gradient == "woodbury":
        # dL = tr(A dS_Z), A = S_Z^-1 - S_Z^-1 S_X S_Z^-1, S_Z = alpha I + beta W_z W_z^T   =>   dL/dZ = sum_k d/dZ < 2 beta A W_z e_k , W_z e_k >.
        # A is applied to the D-vectors W_z e_k through the Woodbury closure: every (beta^-1 I + alpha^-1 G)^-1 solve sits between
        # W_z^T and W_z, which annihilate the Gram's null directions (one per softmax point) where that solve is dominated by the
        # fp32 rounding noise of G at the scale configs.  (The Gram-space form, variational_grad_scalable_exact, applies that inverse
        # to one-hot columns directly and loses accuracy there when alpha / beta is below the Gram's rounding noise.)
        Wz, Sz_inv, S_vp, d_z = parts["Wz"], parts["Sz_inv"], parts["S_vp"], parts["d_z"]
        beta = N / int(Zt.shape[0])
        from . import _cabi as cabi
        blk = _probe_block(D, 256, models=(Wz._lip_model, S_vp._lip_model), zgrad_mode=cabi.ZGRAD_W)
        onehot = torch.eye(d_z, device=Zt.device, dtype=torch.float32)
        dZ = torch.zeros(int(Zt.shape[0]), Wz._lip_model.Z.shape[1], device=Zt.device, dtype=torch.float32)
        for k0 in range(0, d_z, blk):
            E = onehot[k0:min(d_z, k0 + blk)]
            y = Sz_inv(Wz(E.reshape((-1,) + parts["inner_shape"])).reshape(E.shape[0], D))       # S_Z^-1 W_z e_k
            y = y - Sz_inv(S_vp(y))                                                               # A W_z e_k
            dZ += Wz.zgrad(y.mul_(2.0 * beta), E)
        return loss, dZ.reshape(Zt.shape)
    if gradient != "hutchinson":
        raise ValueError(f"gradient must be 'exact', 'woodbury' or 'hutchinson', got {gradient!r}")
    S_vp, Sz_inv = parts["S_vp"], parts["Sz_inv"]
    Sz_vp = compute_curvature_approx(state, Zt, model_type, alpha, full_set_size=N)
    from . import _dist

    def local(E, _):                                      # this rank's probe rows (all of them without a process group)
        a = Sz_inv(E)                                     # S_Z^-1 eps
        b = E - Sz_inv(S_vp(E))                           # eps - S_Z^-1 S_X eps
        return Sz_vp.zgrad(a, b)

    # probes shard over ranks: one all-reduce of the [M, in] gradient
    dZ = _dist.zgrad_sharded(local, probes, probes) / probes.shape[0]
    return loss, dZ.reshape(Zt.shape)


def _scalable_parts(Zt, Xt, state, alpha, model_type, N):
    """The operators the scalable objective and its gradient share: S_vp over the minibatch (:108-110), W_z / W_z^T (:114-116) and
    S_Z^-1 through the Woodbury identity (:127-132) — built once per step."""
    M = int(Zt.shape[0])
    beta = N / M
    S_vp = compute_curvature_approx(state, Xt, model_type, alpha, full_set_size=N)
    Wz, WzT = compute_W_vps(state, Zt, model_type, full_set_size=None)
    bm = Wz._lip_model
    inner_shape = (M,) if model_type == "regressor" else (M, bm.K)
    d_z = M * bm.K
    D = bm.D
    WzTWz = build_WTW(Wz, WzT, inner_shape, d_z, dtype=torch.float32, block=1)                   # :126
    Kmat = torch.eye(d_z, device=WzTWz.device, dtype=torch.float64) / beta + WzTWz.double() / alpha
    LU, piv = torch.linalg.lu_factor(Kmat)

    @matfree.batched
    def Sz_inv(V):   # S_Z^-1 v = v/alpha - alpha^-2 W_z (beta^-1 I + alpha^-1 W_z^T W_z)^-1 W_z^T v
        V = dev_f32(V).reshape(-1, D)
        u = WzT(V).reshape(V.shape[0], d_z)
        x = torch.linalg.lu_solve(LU, piv, u.double().T).T.float()
        return bm.w(x.reshape((V.shape[0],) + inner_shape), scale=Wz._lip_scale, add=V, add_scale=-alpha,
                    batched=True).mul_(-1.0 / alpha ** 2)      # -(1/alpha^2) (Wz x - alpha v) = v/alpha - Wz x / alpha^2

    return dict(S_vp=S_vp, Wz=Wz, WzT=WzT, Sz_inv=Sz_inv, inner_shape=inner_shape, d_z=d_z, WzTWz=WzTWz)


def woodbury_inverse(state, Z, model_type, alpha, full_set_size, X=None):
    """S_Z^-1 as a batched closure (train_inducing.py:127-132)."""
    Zt = dev_f32(Z)
    return _scalable_parts(Zt, Zt if X is None else dev_f32(X), state, float(alpha), model_type, full_set_size)["Sz_inv"]


def optimize_step(Z, X, map_model_state, alpha, opt_state, rng, zoptimizer, num_mc_samples=None, model_type="classifier",
                  full_set_size=None, scalable=True, st_samples=256, slq_samples=2, slq_num_matvecs=None, *, exact=False):
    """train_inducing.py:198-232: one optimiser step on Z.  `zoptimizer` follows the optax protocol the reference uses
    (`update(grads, opt_state, params) -> (updates, new_opt_state)`, updates are ADDED: optax.apply_updates); utils.adam / utils.sgd
    are minimal stand-ins (optax is not in this image).  scalable=False -> variational_grad_dense (:212-222);
    scalable=True -> variational_grad_scalable (stochastic loss, exact gradient; see there) or, with exact=True, the exact-Gram form for both."""
    if not scalable:
        loss, grads = variational_grad_dense(Z, X, map_model_state, alpha, model_type, rng, full_set_size=full_set_size)
    elif exact:
        loss, grads = variational_grad_scalable_exact(Z, X, map_model_state, alpha, model_type, rng, full_set_size=full_set_size)
    else:
        loss, grads = variational_grad_scalable(Z, X, map_model_state, alpha, model_type, rng, full_set_size=full_set_size,
                                                st_samples=st_samples, slq_samples=slq_samples, slq_num_matvecs=slq_num_matvecs)
    updates, new_opt_state = zoptimizer.update(grads, opt_state, dev_f32(Z))
    return dev_f32(Z) + updates, new_opt_state, loss
