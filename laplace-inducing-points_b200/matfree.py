"""B200 replacements for the third-party routines the reference's hot path calls:

  * jax.scipy.sparse.linalg.cg                      -> cg            (stochtrace.py:146,192; sample.py:71)
  * matfree.decomp.tridiag_sym / decomp.bidiag      -> decomp.*      (sample.py:114; train_inducing.py:156)
  * matfree.funm.funm_lanczos_sym / integrand_funm_sym / integrand_funm_product_logdet / dense_funm_*   -> funm.*
  * matfree.stochtrace.estimator / sampler_rademacher / sampler_normal                                  -> stochtrace.*

matfree / jax are not in this image and matfree is unpinned in the reference (requirements.txt:5); the
algorithms are restated from their published form (see oracle/lip_oracle.py header: "parity unpinned").
Every routine is batched over a leading probe axis: vectors are [n] or [B, n].  The loops are host-driven
(matvec closures are Python callables, as in the reference) but all vector work and all scalars stay on the
device: lip_reorth / lip_cg_step / lip_tridiag_funm / lip_basis_combine of include/lip_b200.h.
"""
from __future__ import annotations

import math
from types import SimpleNamespace
from typing import Callable, Optional

import torch

from . import _cabi as cabi
from ._runtime import _require_cuda, dev_f32, ptr, scratch, stream


def _pad4(n: int) -> int:
    return (n + 3) // 4 * 4


def _as2d(v):
    v = dev_f32(v)
    single = v.dim() == 1
    return (v.reshape(1, -1) if single else v.reshape(v.shape[0], -1)), single


def _apply(matvec, V, single):
    """Call a user matvec on [B,n]; closures flagged _lip_batched take the whole batch, others one row at a time."""
    if getattr(matvec, "_lip_batched", False):
        out = matvec(V[0] if single else V)
        out = dev_f32(out)
        return out.reshape(V.shape[0], -1)
    return torch.stack([dev_f32(matvec(V[b])).reshape(-1) for b in range(V.shape[0])])


def batched(fn):
    """Mark a closure as accepting a leading probe axis ([B, n] in -> [B, m] out)."""
    fn._lip_batched = True
    return fn


# ---------------------------------------------------------------------------------------------- CG
def cg(A: Callable, b, x0=None, *, tol=1e-5, atol=0.0, maxiter=None, check_every=4):
    """jax.scipy.sparse.linalg.cg semantics: x0 = 0, stop when r.r <= max(tol^2 b.b, atol^2) or after
    maxiter (default 10 n) iterations, identity preconditioner.  Returns (x, info) with info = iterations [B]."""
    if x0 is not None:
        raise NotImplementedError("cg: x0 is not used by the reference's call sites")
    L = cabi.lib()
    Bm, single = _as2d(b)
    Bm = Bm.contiguous()
    nb, n = Bm.shape
    maxiter = 10 * n if maxiter is None else int(maxiter)
    dev = Bm.device
    x = torch.empty_like(Bm)
    r = torch.empty_like(Bm)
    p = torch.empty_like(Bm)
    gamma = torch.empty(nb, device=dev)
    thresh = torch.empty(nb, device=dev)
    active = torch.empty(nb, device=dev, dtype=torch.int32)
    iters = torch.empty(nb, device=dev, dtype=torch.int32)
    sc, _ = scratch(L.lip_dot_scratch_bytes(n, nb))
    cabi.check(L.lip_cg_init(ptr(Bm), ptr(x), ptr(r), ptr(p), ptr(gamma), ptr(thresh), ptr(active), ptr(iters),
                             float(tol), float(atol), n, nb, sc, stream()), "lip_cg_init")
    k = 0
    while k < maxiter:
        if k % check_every == 0 and not bool(active.any().item()):
            break
        Ap = _apply(A, p, single).contiguous()
        cabi.check(L.lip_cg_step(ptr(x), ptr(r), ptr(p), ptr(Ap), ptr(gamma), ptr(thresh), ptr(active), ptr(iters),
                                 n, nb, sc, stream()), "lip_cg_step")
        k += 1
    return (x[0] if single else x), (iters[0] if single else iters)


# ---------------------------------------------------------------------------------------------- decompositions
class _Tridiag:
    """Result of tridiag_sym: basis Q [B, k, ldq] (rows are Lanczos vectors), diag [B,k], off [B,k-1]."""

    def __init__(self, Q, ldq, n, diag, off):
        self.Q, self.ldq, self.n, self.diag, self.off = Q, ldq, n, diag, off


def _tridiag_sym(num_matvecs: int, *, keep_basis=True):
    """matfree.decomp.tridiag_sym(k), reortho='full' (Arnoldi form, T = (H + H^T)/2 on its three diagonals)."""
    k = int(num_matvecs)

    def decompose(matvec, vec):
        L = cabi.lib()
        V, single = _as2d(vec)
        nb, n = V.shape
        if k > n:
            raise ValueError(f"num_matvecs={k} exceeds the operator dimension {n}")
        dev = V.device
        ldq = _pad4(n)
        Q = torch.zeros(nb, k, ldq, device=dev)
        H = torch.zeros(nb, k, k, device=dev)          # H[b, step, coeff]
        lengths = torch.zeros(nb, k, device=dev)       # |v| after step i
        v = torch.zeros(nb, ldq, device=dev)
        v[:, :n] = V
        qc = torch.empty(nb, n, device=dev)
        length = torch.empty(nb, device=dev)
        hbuf = torch.zeros(nb, k, device=dev)
        sc_dot, _ = scratch(max(L.lip_dot_scratch_bytes(n, nb), L.lip_reorth_scratch_bytes(n, nb, k)))
        cabi.check(L.lip_dot(ptr(v), ptr(v), ptr(length), n, nb, ldq, ldq, sc_dot, stream()))
        length.sqrt_()
        for i in range(k):
            qi = Q[:, i, :]
            # q_i = v / |v|  (stored in the basis and, contiguous, as the matvec input)
            cabi.check(L.lip_scale(ptr(length), 1, ptr(v), C_void(qi), n, nb, ldq, k * ldq, stream()))
            cabi.check(L.lip_scale(ptr(length), 1, ptr(v), ptr(qc), n, nb, ldq, n, stream()))
            w = _apply(matvec, qc, single)
            v[:, :n] = w
            cabi.check(L.lip_reorth(ptr(Q), ldq, i + 1, k, ptr(v), ldq, ptr(hbuf), ptr(length), 2, n, nb,
                                    sc_dot, stream()), "lip_reorth")
            H[:, i, :i + 1] = hbuf[:, :i + 1]
            lengths[:, i] = length
        idx = torch.arange(k, device=dev)
        diag = H[:, idx, idx].contiguous()
        if k > 1:
            upper = H[:, idx[1:], idx[:-1]]            # H[i, i+1] = q_i . A q_{i+1}  (coeff i at step i+1)
            off = (0.5 * (upper + lengths[:, :-1])).contiguous()
        else:
            off = torch.zeros(nb, 0, device=dev)
        return _Tridiag(Q, ldq, n, diag, off), single

    return decompose


def C_void(t: torch.Tensor):
    import ctypes
    return ctypes.c_void_p(t.data_ptr())


class _Bidiag:
    def __init__(self, alphas, betas):
        self.alphas, self.betas = alphas, betas


def _bidiag(num_matvecs: int):
    """matfree.decomp.bidiag(k): Golub-Kahan-Lanczos, full re-orthogonalisation of both bases.
    decompose(Av, vA, v0): vA is A^T (matfree derives it with jax.vjp; closures from this package carry it as
    `._lip_transpose`, otherwise pass it)."""
    k = int(num_matvecs)

    def decompose(Av, vA, v0):
        L = cabi.lib()
        V0, single = _as2d(v0)
        nb, ncols = V0.shape
        dev = V0.device
        ldv = _pad4(ncols)
        vk = torch.zeros(nb, ldv, device=dev)
        vk[:, :ncols] = V0
        tmp = torch.empty(nb, device=dev)
        vc = torch.empty(nb, ncols, device=dev)
        sc0, _ = scratch(L.lip_dot_scratch_bytes(ncols, nb))
        cabi.check(L.lip_dot(ptr(vk), ptr(vk), ptr(tmp), ncols, nb, ldv, ldv, sc0, stream()))
        tmp.sqrt_()
        cabi.check(L.lip_scale(ptr(tmp), 1, ptr(vk), ptr(vk), ncols, nb, ldv, ldv, stream()))
        cabi.check(L.lip_scale(ptr(tmp), 1, ptr(V0.contiguous()), ptr(vc), ncols, nb, ncols, ncols, stream()))
        probe = _apply(Av, vc, single)
        nrows = probe.shape[1]
        ldu = _pad4(nrows)
        if k > min(nrows, ncols):
            raise ValueError(f"num_matvecs={k} exceeds the operator dimensions ({nrows}, {ncols})")
        Us = torch.zeros(nb, k, ldu, device=dev)
        Vs = torch.zeros(nb, k, ldv, device=dev)
        alphas = torch.zeros(nb, k, device=dev)
        betas = torch.zeros(nb, k, device=dev)
        beta = torch.zeros(nb, device=dev)
        alpha = torch.empty(nb, device=dev)
        uk = torch.zeros(nb, ldu, device=dev)
        uc = torch.empty(nb, nrows, device=dev)
        nrm = torch.empty(nb, device=dev)
        ones = torch.ones(nb, device=dev)
        sc, _ = scratch(max(L.lip_reorth_scratch_bytes(max(nrows, ncols), nb, k),
                            L.lip_dot_scratch_bytes(max(nrows, ncols), nb)))
        for i in range(k):
            Vs[:, i, :] = vk
            betas[:, i] = beta
            vc.copy_(vk[:, :ncols])
            Avk = probe if i == 0 else _apply(Av, vc, single)
            uk[:, :nrows] = Avk
            if i > 0:   # uk = A vk - beta * U_{i-1}
                nbeta = -beta
                cabi.check(L.lip_axpby(ptr(nbeta), C_void(Us[:, i - 1, :]), ptr(ones), ptr(uk),
                                       nrows, nb, k * ldu, ldu, stream()))
            # alpha = |uk|; uk /= alpha; CGS against all stored U; renormalise
            cabi.check(L.lip_dot(ptr(uk), ptr(uk), ptr(alpha), nrows, nb, ldu, ldu, sc, stream()))
            alpha.sqrt_()
            cabi.check(L.lip_scale(ptr(alpha), 1, ptr(uk), ptr(uk), nrows, nb, ldu, ldu, stream()))
            cabi.check(L.lip_reorth(ptr(Us), ldu, i, k, ptr(uk), ldu, None, ptr(nrm), 1, nrows, nb, sc, stream()))
            cabi.check(L.lip_scale(ptr(nrm), 1, ptr(uk), ptr(uk), nrows, nb, ldu, ldu, stream()))
            Us[:, i, :] = uk
            alphas[:, i] = alpha
            uc.copy_(uk[:, :nrows])
            w = _apply(vA, uc, single)
            # vk = A^T uk - alpha * V_i
            vnew = torch.zeros(nb, ldv, device=dev)
            vnew[:, :ncols] = w
            nalpha = -alpha
            cabi.check(L.lip_axpby(ptr(nalpha), C_void(Vs[:, i, :]), ptr(ones), ptr(vnew),
                                   ncols, nb, k * ldv, ldv, stream()))
            beta = torch.empty(nb, device=dev)
            cabi.check(L.lip_dot(ptr(vnew), ptr(vnew), ptr(beta), ncols, nb, ldv, ldv, sc, stream()))
            beta.sqrt_()
            cabi.check(L.lip_scale(ptr(beta), 1, ptr(vnew), ptr(vnew), ncols, nb, ldv, ldv, stream()))
            cabi.check(L.lip_reorth(ptr(Vs), ldv, i + 1, k, ptr(vnew), ldv, None, ptr(nrm), 1, ncols, nb, sc, stream()))
            cabi.check(L.lip_scale(ptr(nrm), 1, ptr(vnew), ptr(vnew), ncols, nb, ldv, ldv, stream()))
            vk = vnew
        return _Bidiag(alphas, betas), single

    return decompose


decomp = SimpleNamespace(tridiag_sym=_tridiag_sym, bidiag=_bidiag)


# ---------------------------------------------------------------------------------------------- funm
_FN = {"log": cabi.FN_LOG, "invsqrt": cabi.FN_INVSQRT, "inv": cabi.FN_INV, "identity": cabi.FN_IDENTITY}


class DenseFunm:
    """A symmetric dense matrix function evaluated by the on-device tridiagonal eigensolver
    (lip_tridiag_funm): V f(clip(lambda, clip_min)) V^T.  clip_min=None -> matfree's own eigh;
    clip_min=1.0 -> the reference's patch (matfree_monkeypatch.py:19)."""

    def __init__(self, fn: str, clip_min: Optional[float] = None):
        if fn not in _FN:
            raise ValueError(f"unknown matrix function {fn!r}; supported {sorted(_FN)}")
        self.fn, self.clip_min = fn, clip_min

    def _run(self, diag, off, want_quad, want_fe1):
        L = cabi.lib()
        nb, k = diag.shape
        quad = torch.empty(nb, device=diag.device) if want_quad else None
        fe1 = torch.empty(nb, k, device=diag.device) if want_fe1 else None
        sc, _ = scratch(L.lip_tridiag_scratch_bytes(k, nb, 1 if want_fe1 else 0))
        cabi.check(L.lip_tridiag_funm(ptr(diag), ptr(off), k, nb, _FN[self.fn],
                                      -1.0 if self.clip_min is None else float(self.clip_min),
                                      ptr(quad), ptr(fe1), None, sc, stream()), "lip_tridiag_funm")
        return quad, fe1

    def quad_e1(self, diag, off):
        return self._run(diag.contiguous(), off.contiguous(), True, False)[0]

    def apply_e1(self, diag, off):
        return self._run(diag.contiguous(), off.contiguous(), False, True)[1]


def _matfun_name(matfun) -> str:
    """Map the callables the reference passes (np.log, lambda x: 1/sqrt(x)) onto device functions by probing."""
    if isinstance(matfun, str):
        return matfun
    x = torch.tensor([4.0], dtype=torch.float64)
    try:
        y = float(matfun(x)[0])
    except Exception:
        import numpy as np
        y = float(matfun(np.array([4.0]))[0])
    for name, val in (("log", math.log(4.0)), ("invsqrt", 0.5), ("inv", 0.25), ("identity", 4.0)):
        if abs(y - val) < 1e-9:
            return name
    raise ValueError("unsupported matrix function for the on-device eigensolver (supported: log, 1/sqrt, 1/x, x)")


def dense_funm_sym_eigh(matfun) -> DenseFunm:
    """matfree.funm.dense_funm_sym_eigh (UNPATCHED: no eigenvalue clip) — tests/test_sample.py:9,337."""
    return DenseFunm(_matfun_name(matfun), None)


def funm_lanczos_sym(dense_funm: DenseFunm, tridiag_sym):
    """matfree.funm.funm_lanczos_sym: f(A) v ~= |v| Q f(T) e1   (sample.py:115)."""

    def estimate(matvec, vec):
        L = cabi.lib()
        V, single = _as2d(vec)
        nb, n = V.shape
        length = torch.linalg.vector_norm(V, dim=1)
        res, _ = tridiag_sym(matvec, V / length[:, None])
        fe1 = dense_funm.apply_e1(res.diag, res.off)
        k = res.diag.shape[1]
        out = torch.empty(nb, res.ldq, device=V.device)
        cabi.check(L.lip_basis_combine(ptr(res.Q), res.ldq, k, k, ptr(fe1), k, ptr(out), res.ldq, n, nb, stream()),
                   "lip_basis_combine")
        out = out[:, :n] * length[:, None]
        return out[0] if single else out

    return batched(estimate)


def integrand_funm_sym(dense_funm: DenseFunm, tridiag_sym):
    """matfree.funm.integrand_funm_sym: v -> |v|^2 e1^T f(T) e1."""

    def quadform(matvec, v0):
        V, single = _as2d(v0)
        length = torch.linalg.vector_norm(V, dim=1)
        res, _ = tridiag_sym(matvec, V / length[:, None])
        q = dense_funm.quad_e1(res.diag, res.off) * length ** 2
        return q[0] if single else q

    return batched(quadform)


def integrand_funm_sym_logdet(tridiag_sym):
    """matfree.funm.integrand_funm_sym_logdet (UNPATCHED) — tests/test_variational.py:5,73-77."""
    return integrand_funm_sym(DenseFunm("log", None), tridiag_sym)


def integrand_funm_product_logdet(bidiag):
    """matfree.funm.integrand_funm_product_logdet (train_inducing.py:157): |v|^2 e1^T V log(S^2) V^T e1 for the GKL
    bidiagonal B = U S V^T, evaluated through the eigen-decomposition of T = B^T B (no clip: matfree's own)."""
    dense = DenseFunm("log", None)

    def quadform(Av, v0, vA=None):
        L = cabi.lib()
        vA_ = vA if vA is not None else getattr(Av, "_lip_transpose", None)
        if vA_ is None:
            raise ValueError("integrand_funm_product_logdet: the transpose operator is required (pass vA=... or use a "
                             "closure built by this package)")
        V, single = _as2d(v0)
        length = torch.linalg.vector_norm(V, dim=1)
        res, _ = bidiag(Av, vA_, V / length[:, None])
        nb, k = res.alphas.shape
        td = torch.empty(nb, k, device=V.device)
        to = torch.empty(nb, max(k - 1, 1), device=V.device)
        cabi.check(L.lip_bidiag_to_tridiag(ptr(res.alphas.contiguous()), ptr(res.betas.contiguous()), ptr(td), ptr(to), k,
                                           nb, stream()))
        q = dense.quad_e1(td, to[:, :k - 1].contiguous() if k > 1 else to) * length ** 2
        return q[0] if single else q

    return batched(quadform)


funm = SimpleNamespace(dense_funm_sym_eigh=dense_funm_sym_eigh, funm_lanczos_sym=funm_lanczos_sym,
                       integrand_funm_sym=integrand_funm_sym, integrand_funm_sym_logdet=integrand_funm_sym_logdet,
                       integrand_funm_product_logdet=integrand_funm_product_logdet)


# ---------------------------------------------------------------------------------------------- stochtrace
def _generator(key):
    if isinstance(key, torch.Generator):
        return key
    g = torch.Generator(device=_require_cuda())
    g.manual_seed(int(key) if key is not None else 0)
    return g


def sampler_rademacher(x_like, /, *, num):
    n = int(dev_f32(x_like).numel())

    def sample(key):
        g = _generator(key)
        return (torch.randint(0, 2, (num, n), generator=g, device=g.device, dtype=torch.int8).float() * 2 - 1)

    return sample


def sampler_normal(x_like, /, *, num):
    n = int(dev_f32(x_like).numel())

    def sample(key):
        g = _generator(key)
        return torch.randn(num, n, generator=g, device=g.device)

    return sample


def estimator(integrand, /, sampler):
    """matfree.stochtrace.estimator: mean over sampler(key) rows of integrand(matvec, row)."""

    def estimate(matvecs, key, *parameters):
        samples = sampler(key)
        if getattr(integrand, "_lip_batched", False):
            return integrand(matvecs, samples, *parameters).mean()
        return torch.stack([integrand(matvecs, s, *parameters) for s in samples]).mean()

    return estimate


stochtrace = SimpleNamespace(estimator=estimator, sampler_rademacher=sampler_rademacher, sampler_normal=sampler_normal)
