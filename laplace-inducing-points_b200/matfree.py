"""B200 replacements for the third-party routines the reference's hot path calls:

  * jax.scipy.sparse.linalg.cg                      -> cg            (stochtrace.py:146,192; sample.py:71)
  * matfree.decomp.tridiag_sym / decomp.bidiag      -> decomp.*      (sample.py:114; train_inducing.py:156)
  * matfree.funm.funm_lanczos_sym / integrand_funm_sym / integrand_funm_product_logdet / dense_funm_*   -> funm.*
  * matfree.stochtrace.estimator / sampler_rademacher / sampler_normal                                  -> stochtrace.*

matfree / jax are not in this image and matfree is unpinned in the reference (requirements.txt:5); the
algorithms are restated from their published form (see oracle/lip_oracle.py header: "parity unpinned").
Every routine is batched over a leading probe axis: vectors are [n] or [B, n].  Each recurrence is ONE native call
(lip_lanczos_tridiag / lip_gkl_bidiag / lip_cg_solve of include/lip_b200.h, csrc/lip_krylov.cu): closures built by this
package (curvature_vp, gkl_target, dense_sym_operator) run without Python in the loop; any other callable is invoked
through the library's mat-vec callback, once per step.  All vector work and all scalars stay on the device.
"""
from __future__ import annotations

import math
from types import SimpleNamespace
from typing import Callable, Optional

import torch

from . import _cabi as cabi
from ._runtime import _require_cuda, dev_f32, ptr, stream


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


# ---------------------------------------------------------------------------------------------- operators
class _NativeOp:
    """A lip_linop (include/lip_b200.h) for a mat-vec closure, plus everything that must outlive the native call.

    Closures built by this package carry what the library needs to run them without Python in the loop:
      _lip_kind == "GGN"       (lla.compute_curvature_approx / ggn.compute_ggn_vp)   -> LIP_LINOP_GGN
      _lip_kind == "GKL"       (gkl_target below)                                   -> LIP_LINOP_GKL
      _lip_kind == "DENSE_SYM" (dense_sym_operator below)                           -> LIP_LINOP_DENSE_SYM
    anything else is wrapped as LIP_LINOP_CALLBACK: the recurrence still runs inside ONE native call and the library calls back
    into Python for the mat-vec only."""

    def __init__(self, Av, vA, B, n_in, n_out, single, symmetric, model=None):
        self.struct = cabi.LinOp()
        self.exc = None
        self.keep = []
        kind = getattr(Av, "_lip_kind", None)
        native = getattr(Av, "_lip_native", True) and getattr(Av, "_lip_model", None) is not None
        if kind == "GGN" and native and symmetric:
            bm = model if model is not None else Av._lip_model
            self.struct.kind, self.struct.model = cabi.LINOP_GGN, bm._h
            self.struct.scale, self.struct.alpha = float(Av._lip_recal), float(Av._lip_alpha)
            self.keep.append(bm)
        elif kind == "GKL" and native and not symmetric:
            bm = model if model is not None else Av._lip_model
            self.struct.kind, self.struct.model = cabi.LINOP_GKL, bm._h
            self.struct.scale, self.struct.alpha = float(Av._lip_scale), float(Av._lip_alpha)
            self.keep.append(bm)
        elif kind == "DENSE_SYM" and symmetric:
            G = Av._lip_dense
            self.struct.kind, self.struct.dense, self.struct.n = cabi.LINOP_DENSE_SYM, G.data_ptr(), int(G.shape[0])
            self.struct.alpha, self.struct.beta = float(Av._lip_alpha), float(Av._lip_beta)
            self.keep.append(G)
        else:
            dev = _require_cuda()
            self.cb_in = torch.empty(B, n_in, device=dev)
            self.cb_out = torch.empty(B, n_out, device=dev)
            self.struct.kind, self.struct.symmetric = cabi.LINOP_CALLBACK, 1 if symmetric else 0
            self.struct.n, self.struct.n_out = n_in, n_out
            self.struct.cb_in, self.struct.cb_out = self.cb_in.data_ptr(), self.cb_out.data_ptr()
            if not symmetric:
                self.cb_in_t = torch.empty(B, n_out, device=dev)
                self.cb_out_t = torch.empty(B, n_in, device=dev)
                self.struct.cb_in_t, self.struct.cb_out_t = self.cb_in_t.data_ptr(), self.cb_out_t.data_ptr()

            def call(_ctx, transpose, nb, _stream):
                try:
                    if transpose:
                        self.cb_out_t[:nb].copy_(_apply(vA, self.cb_in_t[:nb], single))
                    else:
                        self.cb_out[:nb].copy_(_apply(Av, self.cb_in[:nb], single))
                    return 0
                except BaseException as e:          # noqa: BLE001  (re-raised by check() after the native call returns)
                    self.exc = e
                    return 1

            self.fn = cabi.MATVEC_FN(call)
            self.struct.fn = self.fn

    def ref(self):
        import ctypes
        return ctypes.byref(self.struct)

    def check(self, rc, what):
        if self.exc is not None:
            exc, self.exc = self.exc, None
            raise exc
        cabi.check(rc, what)

    def workspace(self, routine, k, B):
        need = cabi.lib().lip_krylov_workspace_bytes(self.ref(), routine, k, B)
        if need == 0:
            raise ValueError("lip_krylov_workspace_bytes: " + cabi.lib().lip_last_error().decode("utf-8", "replace"))
        buf = torch.empty(need, dtype=torch.uint8, device=_require_cuda())   # local: lives exactly as long as the recurrence
        return buf, need


def gkl_target(WTfun, Wfun, alpha):
    """bidiag_target of train_inducing.py:166-169 for closures from ggn.compute_W_vps:  v -> [sqrt(alpha) v ; W^T v]  in R^{D+d}
    (its transpose, which matfree obtains with jax.vjp, is `._lip_transpose`).  Passed to decomp.bidiag / the product-logdet
    integrand the whole Golub-Kahan recurrence runs natively (LIP_LINOP_GKL)."""
    bm = WTfun._lip_model
    D, d = bm.D, bm.M * bm.K
    sa = math.sqrt(float(alpha))

    def Av(v):
        V, single = _as2d(v)
        out = torch.cat([sa * V, WTfun(V).reshape(V.shape[0], d)], dim=1)
        return out[0] if single else out

    def vA(u):
        U, single = _as2d(u)
        out = Wfun(U[:, D:].reshape((-1,) + ((bm.M,) if bm.model_type == "regressor" else (bm.M, bm.K)))).reshape(U.shape[0], D)
        out = out + sa * U[:, :D]
        return out[0] if single else out

    for f, t in ((Av, vA), (vA, Av)):
        f._lip_batched = True
        f._lip_transpose = t
    Av._lip_kind, Av._lip_model, Av._lip_scale, Av._lip_alpha = "GKL", bm, WTfun._lip_scale, float(alpha)
    return Av


def dense_sym_operator(G, alpha, beta):
    """u -> alpha u + beta G u for a dense symmetric G (inner_fun_flat of sample.py:120-125); runs as LIP_LINOP_DENSE_SYM."""
    G = dev_f32(G)
    if G.dim() != 2 or G.shape[0] != G.shape[1]:
        raise ValueError(f"dense_sym_operator: square matrix expected, got {tuple(G.shape)}")

    def mv(u):
        U, single = _as2d(u)
        U = U.contiguous()
        op = _NativeOp(mv, None, U.shape[0], G.shape[0], G.shape[0], single, symmetric=True)
        out = torch.empty_like(U)
        ws, need = op.workspace(cabi.KRYLOV_APPLY, 1, U.shape[0])
        op.check(cabi.lib().lip_linop_apply(op.ref(), ptr(U), ptr(out), U.shape[0], 0, ptr(ws), need, stream()), "lip_linop_apply")
        return out[0] if single else out

    mv._lip_batched, mv._lip_kind, mv._lip_dense, mv._lip_alpha, mv._lip_beta = True, "DENSE_SYM", G, float(alpha), float(beta)
    return mv


# ---------------------------------------------------------------------------------------------- CG
def cg(A: Callable, b, x0=None, *, tol=1e-5, atol=0.0, maxiter=None, check_every=8):
    """jax.scipy.sparse.linalg.cg semantics: x0 = 0, stop when r.r <= max(tol^2 b.b, atol^2) or after
    maxiter (default 10 n) iterations, identity preconditioner.  Returns (x, info) with info = iterations [B].
    One native call (lip_cg_solve); check_every = 0 never synchronises (exactly maxiter masked iterations)."""
    if x0 is not None:
        raise NotImplementedError("cg: x0 is not used by the reference's call sites")
    L = cabi.lib()
    Bm, single = _as2d(b)
    Bm = Bm.contiguous()
    nb, n = Bm.shape
    op = _NativeOp(A, None, nb, n, n, single, symmetric=True)
    x = torch.empty_like(Bm)
    iters = torch.empty(nb, device=Bm.device, dtype=torch.int32)
    ws, need = op.workspace(cabi.KRYLOV_CG, 1, nb)
    rc = L.lip_cg_solve(op.ref(), ptr(Bm), ptr(x), nb, float(tol), float(atol), -1 if maxiter is None else int(maxiter),
                        int(check_every), ptr(iters), ptr(ws), need, stream())
    op.check(rc, "lip_cg_solve")
    return (x[0] if single else x), (iters[0] if single else iters)


# ---------------------------------------------------------------------------------------------- decompositions
class _Tridiag:
    """Result of tridiag_sym: basis Q [B, k, ldq] (rows are Lanczos vectors), diag [B,k], off [B,k-1], norm0 [B] = |v0|."""

    def __init__(self, Q, ldq, n, diag, off, norm0=None):
        self.Q, self.ldq, self.n, self.diag, self.off, self.norm0 = Q, ldq, n, diag, off, norm0


def _tridiag_sym(num_matvecs: int, *, keep_basis=True, passes=2):
    """matfree.decomp.tridiag_sym(k), reortho='full' (Arnoldi form, T = (H + H^T)/2 on its three diagonals): one
    lip_lanczos_tridiag call."""
    k = int(num_matvecs)

    def decompose(matvec, vec):
        L = cabi.lib()
        V, single = _as2d(vec)
        V = V.contiguous()
        nb, n = V.shape
        if k > n:
            raise ValueError(f"num_matvecs={k} exceeds the operator dimension {n}")
        dev = V.device
        ldq = _pad4(n)
        op = _NativeOp(matvec, None, nb, n, n, single, symmetric=True)
        Q = torch.empty(nb, k, ldq, device=dev)
        diag = torch.empty(nb, k, device=dev)
        off = torch.empty(nb, max(k - 1, 1), device=dev)
        norm0 = torch.empty(nb, device=dev)
        ws, need = op.workspace(cabi.KRYLOV_LANCZOS, k, nb)
        rc = L.lip_lanczos_tridiag(op.ref(), ptr(V), n, k, nb, int(passes), ptr(Q), ldq, ptr(diag), ptr(off), ptr(norm0),
                                   ptr(ws), need, stream())
        op.check(rc, "lip_lanczos_tridiag")
        return _Tridiag(Q, ldq, n, diag, off[:, :k - 1].contiguous(), norm0), single

    return decompose


class _Bidiag:
    def __init__(self, alphas, betas, Us=None, Vs=None, norm0=None):
        self.alphas, self.betas, self.Us, self.Vs, self.norm0 = alphas, betas, Us, Vs, norm0


def _bidiag(num_matvecs: int):
    """matfree.decomp.bidiag(k): Golub-Kahan-Lanczos, full re-orthogonalisation of both bases: one lip_gkl_bidiag call.
    decompose(Av, vA, v0): vA is A^T (matfree derives it with jax.vjp; closures from this package carry it as
    `._lip_transpose`, otherwise pass it)."""
    k = int(num_matvecs)

    def decompose(Av, vA, v0):
        L = cabi.lib()
        V0, single = _as2d(v0)
        V0 = V0.contiguous()
        nb, ncols = V0.shape
        dev = V0.device
        if getattr(Av, "_lip_kind", None) == "GKL" and getattr(Av, "_lip_model", None) is not None:
            bm = Av._lip_model
            nrows = bm.D + bm.M * bm.K
        elif hasattr(Av, "_lip_out_dim"):
            nrows = int(Av._lip_out_dim)
        else:                      # learn the row count from one application (matfree traces the function instead)
            nrows = int(_apply(Av, V0[:1], single).shape[1])
        if k > min(nrows, ncols):
            raise ValueError(f"num_matvecs={k} exceeds the operator dimensions ({nrows}, {ncols})")
        op = _NativeOp(Av, vA, nb, ncols, nrows, single, symmetric=False)
        ldu, ldv = _pad4(nrows), _pad4(ncols)
        Us = torch.empty(nb, k, ldu, device=dev)
        Vs = torch.empty(nb, k, ldv, device=dev)
        alphas = torch.empty(nb, k, device=dev)
        betas = torch.empty(nb, k, device=dev)
        norm0 = torch.empty(nb, device=dev)
        ws, need = op.workspace(cabi.KRYLOV_GKL, k, nb)
        rc = L.lip_gkl_bidiag(op.ref(), ptr(V0), ncols, k, nb, ptr(Us), ldu, ptr(Vs), ldv, ptr(alphas), ptr(betas), ptr(norm0),
                              ptr(ws), need, stream())
        op.check(rc, "lip_gkl_bidiag")
        return _Bidiag(alphas, betas, Us, Vs, norm0), single

    decompose._lip_num_matvecs = k
    return decompose


decomp = SimpleNamespace(tridiag_sym=_tridiag_sym, bidiag=_bidiag)


# ---------------------------------------------------------------------------------------------- funm
_FN = {"log": cabi.FN_LOG, "invsqrt": cabi.FN_INVSQRT, "inv": cabi.FN_INV, "identity": cabi.FN_IDENTITY,
       "sampler": cabi.FN_SAMPLER}


class DenseFunm:
    """A symmetric dense matrix function evaluated by the on-device tridiagonal eigensolver
    (lip_tridiag_funm): V f(clip(lambda, clip_min)) V^T.  clip_min=None -> matfree's own eigh;
    clip_min=1.0 -> the reference's patch (matfree_monkeypatch.py:19)."""

    def __init__(self, fn: str, clip_min: Optional[float] = None, params=None):
        if fn not in _FN:
            raise ValueError(f"unknown matrix function {fn!r}; supported {sorted(_FN)}")
        self.fn, self.clip_min = fn, clip_min
        self.params = None if params is None else tuple(float(p) for p in params)   # "sampler": (alpha, beta, tau)

    def _run(self, diag, off, want_quad, want_fe1):
        import ctypes
        L = cabi.lib()
        nb, k = diag.shape
        quad = torch.empty(nb, device=diag.device) if want_quad else None
        fe1 = torch.empty(nb, k, device=diag.device) if want_fe1 else None
        sc = torch.empty(L.lip_tridiag_scratch_bytes(k, nb, 1 if want_fe1 else 0), dtype=torch.uint8, device=diag.device)
        par = (ctypes.c_float * 3)(*self.params) if self.params is not None else None
        cabi.check(L.lip_tridiag_funm_p(ptr(diag), ptr(off), k, nb, _FN[self.fn],
                                        -1.0 if self.clip_min is None else float(self.clip_min), par,
                                        ptr(quad), ptr(fe1), None, ptr(sc), stream()), "lip_tridiag_funm")
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
        res, _ = tridiag_sym(matvec, V)                       # normalises v itself; res.norm0 = |v|
        fe1 = dense_funm.apply_e1(res.diag, res.off) * res.norm0[:, None]
        k = res.diag.shape[1]
        out = torch.empty(nb, res.ldq, device=V.device)
        cabi.check(L.lip_basis_combine(ptr(res.Q), res.ldq, k, k, ptr(fe1.contiguous()), k, ptr(out), res.ldq, n, nb, stream()),
                   "lip_basis_combine")
        out = out[:, :n]
        return out[0] if single else out

    return batched(estimate)


def integrand_funm_sym(dense_funm: DenseFunm, tridiag_sym):
    """matfree.funm.integrand_funm_sym: v -> |v|^2 e1^T f(T) e1."""

    def quadform(matvec, v0):
        V, single = _as2d(v0)
        res, _ = tridiag_sym(matvec, V)
        q = dense_funm.quad_e1(res.diag, res.off) * res.norm0 ** 2
        return q[0] if single else q

    return batched(quadform)


def integrand_funm_sym_logdet(tridiag_sym):
    """matfree.funm.integrand_funm_sym_logdet (UNPATCHED) — tests/test_variational.py:5,73-77."""
    return integrand_funm_sym(DenseFunm("log", None), tridiag_sym)


def integrand_funm_product_logdet(bidiag):
    """matfree.funm.integrand_funm_product_logdet (train_inducing.py:157): |v|^2 e1^T V log(S^2) V^T e1 for the GKL
    bidiagonal B = U S V^T, evaluated through the eigen-decomposition of T = B^T B.  matfree takes the SVD of B, whose squared
    singular values are >= 0 by construction; the eigenvalue route can return -1e-13 |T| for a zero singular value of the
    decoupled post-breakdown block (weight ~1e-13, but its log is NaN), so the spectrum is floored at 1e-30."""
    dense = DenseFunm("log", 1e-30)

    def quadform(Av, v0, vA=None):
        L = cabi.lib()
        vA_ = vA if vA is not None else getattr(Av, "_lip_transpose", None)
        if vA_ is None:
            raise ValueError("integrand_funm_product_logdet: the transpose operator is required (pass vA=... or use a "
                             "closure built by this package)")
        V, single = _as2d(v0)
        k_native = getattr(bidiag, "_lip_num_matvecs", None)
        if (k_native is not None and getattr(Av, "_lip_kind", None) == "GKL" and getattr(Av, "_lip_model", None) is not None
                and vA_ is getattr(Av, "_lip_transpose", None)):
            # this package's own bidiag_target closure through this package's own decomposition: the whole integrand is ONE native
            # call (lip_slq_quadrature) that never materialises the u basis (csrc/lip_krylov.cu gkl_run, reduced form)
            q = slq_quadrature(Av, V, k_native, form="gkl", fn="log", clip_min=None)
            return q[0] if single else q
        res, _ = bidiag(Av, vA_, V)
        length = res.norm0
        nb, k = res.alphas.shape
        td = torch.empty(nb, k, device=V.device)
        to = torch.empty(nb, max(k - 1, 1), device=V.device)
        cabi.check(L.lip_bidiag_to_tridiag(ptr(res.alphas.contiguous()), ptr(res.betas.contiguous()), ptr(td), ptr(to), k,
                                           nb, stream()))
        q = dense.quad_e1(td, to[:, :k - 1].contiguous() if k > 1 else to) * length ** 2
        return q[0] if single else q

    return batched(quadform)


def slq_quadrature(matvec, probes, num_matvecs, *, form="gkl", fn="log", clip_min=None, comm=None, model=None):
    """The fused native SLQ integrand (lip_slq_quadrature / lip_slq_quadrature_sharded): per-probe |v|^2 e1^T f(T) e1, [B].
    form "gkl": matvec is a gkl_target closure (integrand_funm_product_logdet, train_inducing.py:156-157);
    form "lanczos": a symmetric closure (integrand_funm_sym; fn="log", clip_min=1.0 is the patched integrand_funm_sym_logdet).
    comm: a _dist.NativeComm whose ranks share the Krylov bases column-wise (every rank passes the same probes).
    model: a BoundModel.clone() to run on instead of the closure's own handle (a second concurrent stream needs its own)."""
    L = cabi.lib()
    P, single = _as2d(probes)
    P = P.contiguous()
    nb, n = P.shape
    sym = form == "lanczos"
    if not sym and form != "gkl":
        raise ValueError(f"slq_quadrature: unknown form {form!r}")
    if sym:
        nout = n
    else:
        bm = getattr(matvec, "_lip_model", None)
        if bm is None or getattr(matvec, "_lip_kind", None) != "GKL":
            raise ValueError("slq_quadrature(form='gkl') needs a matfree.gkl_target closure")
        nout = bm.D + bm.M * bm.K
    op = _NativeOp(matvec, getattr(matvec, "_lip_transpose", None), nb, n, nout, single, symmetric=sym, model=model)
    world = 1 if comm is None else comm.world
    if world > 1 and op.struct.kind not in (cabi.LINOP_GGN, cabi.LINOP_GKL):
        raise ValueError("sharded slq_quadrature needs one of this package's model closures (curvature_vp / gkl_target)")
    f = cabi.SLQ_LANCZOS if sym else cabi.SLQ_GKL
    need = L.lip_slq_workspace_bytes(op.ref(), f, int(num_matvecs), nb, world)
    if need == 0:
        raise ValueError("lip_slq_workspace_bytes: " + L.lip_last_error().decode("utf-8", "replace"))
    ws = torch.empty(need, dtype=torch.uint8, device=P.device)
    out = torch.empty(nb, device=P.device)
    rc = L.lip_slq_quadrature_sharded(op.ref(), None if comm is None else comm.handle, ptr(P), n, int(num_matvecs), nb, f, _FN[fn],
                                      -1.0 if clip_min is None else float(clip_min), ptr(out), ptr(ws), need, stream())
    op.check(rc, "lip_slq_quadrature")
    return out[0] if single else out


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
        from ._runtime import exact_tf32_block
        g = _generator(key)
        out = exact_tf32_block(num, n, g.device)
        out.copy_(torch.randint(0, 2, (num, n), generator=g, device=g.device, dtype=torch.int8).float() * 2 - 1)
        return out

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
