"""Drop-in mirror of /root/reference/src/stochtrace.py on the B200 path.

Same names and argument meaning.  `seed` is an int / torch.Generator (JAX's threefry keys cannot be reproduced
without JAX; SURVEY §2.1: probes are inputs) and every estimator also takes `eps=` to supply the probe matrix
explicitly, which is how the parity tests feed identical probes to this path and to the oracle.
Probe batching: where the reference vmaps `Xfun` over probe rows, closures flagged `_lip_batched` receive the
whole [B, n] block in one call.
"""
from __future__ import annotations

import torch

from ._runtime import _require_cuda, dev_f32, exact_tf32_block, is_exact_tf32
from .matfree import _generator, cg


def vmap(fn, in_axes=0, out_axes=0):
    """jax.vmap for matvec closures: batched closures get the whole block, others are looped."""

    def mapped(X):
        if in_axes == 0 and out_axes == 0 and is_exact_tf32(X) and getattr(fn, "_lip_batched", False):
            return fn(X)                        # padded +-1 probe rows go to the operator as they are (read in place by the GEMMs)
        X = dev_f32(X)
        Xr = X if in_axes == 0 else X.transpose(0, 1)
        if getattr(fn, "_lip_batched", False):
            Y = fn(Xr.contiguous())
        else:
            Y = torch.stack([dev_f32(fn(x)) for x in Xr])
        return Y if out_axes == 0 else Y.transpose(0, 1)

    return mapped


def pack_rademacher(eps) -> "np.ndarray":
    """+-1 probe matrix [B, n] (host) -> packed bits [B, ceil(n/8)] uint8 (numpy.packbits order): the wire format for
    host-resident Rademacher probes (1 bit per element over PCIe)."""
    import numpy as np
    e = np.asarray(eps.cpu() if isinstance(eps, torch.Tensor) else eps)
    return np.packbits(e > 0, axis=1)


def unpack_rademacher(bits, n: int, out=None):
    """Packed bits [B, ceil(n/8)] (a uint8 CUDA tensor) -> float32 +-1 probes [B, n] on the device (lip_unpack_rademacher)."""
    from . import _cabi as cabi
    from ._runtime import ptr, stream
    if not (isinstance(bits, torch.Tensor) and bits.is_cuda and bits.dtype == torch.uint8 and bits.dim() == 2):
        raise ValueError("unpack_rademacher: bits must be a 2-D uint8 CUDA tensor")
    bits = bits.contiguous()
    B = bits.shape[0]
    if out is None:
        out = exact_tf32_block(B, n, bits.device)         # padded rows + the exactly-TF32 mark: lip_ggn_vp_ex reads them in place
    if out.dim() != 2 or out.shape != (B, n) or out.stride(1) != 1 or out.dtype != torch.float32:
        raise ValueError("unpack_rademacher: out must be a [B, n] float32 tensor with unit column stride")
    cabi.check(cabi.lib().lip_unpack_rademacher_ld(ptr(bits), bits.shape[1], ptr(out), out.stride(0), n, B, stream()),
               "lip_unpack_rademacher")
    out._lip_exact_tf32 = True                            # +-1 by construction
    return out


def _rademacher(seed, shape):
    """+-1 probes [B, n] in padded rows, marked exactly-TF32 (src/stochtrace.py:28): the layout lip_ggn_vp_ex reads in place."""
    g = _generator(seed)
    out = exact_tf32_block(shape[0], shape[1], g.device)
    out.copy_(torch.randint(0, 2, tuple(shape), generator=g, device=g.device, dtype=torch.int8).float() * 2 - 1)
    return out


def mark_exact_tf32(t):
    """Declare that a [B, n] CUDA float32 tensor holds only values exactly representable in TF32 (+-1 probes, one-hot rows, small
    integers): operators may then skip the TF32 split of the block (LIP_PROBES_EXACT_TF32).  The promise is the caller's; it does
    not survive slicing or arithmetic (the attribute lives on this tensor object), and it must be dropped if the data is edited."""
    t._lip_exact_tf32 = True
    return t


def _normal(seed, shape):
    g = _generator(seed)
    return torch.randn(*shape, generator=g, device=g.device)


def stochastic_trace_estimator_dense(X, seed, num_samples=1_000, *, eps=None):
    """stochtrace.py:7-19"""
    X = dev_f32(X)
    Eps = dev_f32(eps) if eps is not None else _rademacher(seed, (num_samples, X.shape[0]))
    return ((Eps @ X.T) * Eps).sum(1).mean()


def stochastic_trace_estimator_mvp(Xfun, D, seed, num_samples=1_000, dtype=torch.float32, *, eps=None):
    """stochtrace.py:22-34: mean_b eps_b . X eps_b with Rademacher probes."""
    Eps = dev_f32(eps) if eps is not None else _rademacher(seed, (num_samples, D))
    Y = vmap(Xfun)(Eps)
    return (Eps * Y).sum(1).mean()


def _project_out(Q, Mx):
    """(I - Q Q^T) Mx without the n x n projector of stochtrace.py:47,65,96."""
    return Mx - Q @ (Q.T @ Mx)


def hutchpp_dense(X, seed, num_samples=10, *, eps=None):
    """stochtrace.py:37-49"""
    X = dev_f32(X)
    e = dev_f32(eps) if eps is not None else _normal(seed, (num_samples * 2, X.shape[0]))
    ns = e.shape[0] // 2
    S, G = e[:ns], e[ns:]
    Q, _ = torch.linalg.qr(X @ S.T)
    PG = _project_out(Q, G.T)
    return torch.trace(Q.T @ X @ Q) + torch.trace(PG.T @ X @ PG) / ns


def hutchpp_mvp(Xfun, D, seed, num_samples=10, *, eps=None):
    """stochtrace.py:52-79; Xfun maps a MATRIX [D,k] -> [D,k] (:64,74)."""
    e = dev_f32(eps) if eps is not None else _normal(seed, (num_samples * 2, D))
    ns = e.shape[0] // 2
    S, G = e[:ns], e[ns:]
    Q, _ = torch.linalg.qr(dev_f32(Xfun(S.T.contiguous())))

    def quad_term(Mx):
        return Mx.T @ dev_f32(Xfun(Mx.contiguous()))

    return torch.trace(quad_term(Q)) + torch.trace(quad_term(_project_out(Q, G.T))) / ns


def hutchpp(Xfun, sampler):
    """stochtrace.py:82-111; Xfun maps a vector; NB the residual term divides by the FULL probe count (:84,109)."""
    e = dev_f32(sampler(...))
    num_samples = e.shape[0]
    S, G = e[:num_samples // 2], e[num_samples // 2:]
    Q, _ = torch.linalg.qr(vmap(Xfun, in_axes=0, out_axes=1)(S), mode="reduced")

    def quad_term(Mx):
        Y = vmap(Xfun, in_axes=1, out_axes=1)(Mx)
        return Mx.T @ Y

    return torch.trace(quad_term(Q)) + torch.trace(quad_term(_project_out(Q, G.T))) / num_samples


def apply_X(Xfun, Mx):
    """stochtrace.py:113-114: rows of Mx are probes; result columns are X @ probe."""
    return vmap(Xfun, in_axes=0, out_axes=1)(Mx)


def hutchpp_v2(Xfun, sampler, *, s1, s2):
    """stochtrace.py:118-135: tr(Q^T X Q) + tr(G_perp X G_perp^T)/s2 with Q = orth(X S^T), G_perp = G - (G Q) Q^T.
    One native call (lip_hutchpp_v2): the [n, s1] QR is a three-pass shifted CholeskyQR on the library's own kernels (float64
    tall-skinny Gram, one-CTA Cholesky, GEMM apply) and the deflation one GEMM; Xfun runs natively when it is one of this package's
    closures, else through the mat-vec callback (2 s1 + s2 products in three batched calls)."""
    from . import _cabi as cabi
    from ._runtime import ptr, stream
    from .matfree import _NativeOp
    e = dev_f32(sampler(...))
    if e.dim() != 2 or e.shape[0] < s1 + s2:
        raise ValueError(f"hutchpp_v2: the sampler returned {tuple(e.shape)}, need at least s1 + s2 = {s1 + s2} probe rows")
    e = e[:s1 + s2].contiguous()
    n = e.shape[1]
    sm = max(int(s1), int(s2))
    op = _NativeOp(Xfun, None, sm, n, n, False, symmetric=True)
    out = torch.empty(1, device=e.device)
    info = torch.zeros(1, device=e.device, dtype=torch.int32)
    L = cabi.lib()
    need = L.lip_krylov_workspace_bytes(op.ref(), cabi.KRYLOV_HUTCHPP, int(s1), int(s2))
    if need == 0:
        raise ValueError("lip_hutchpp_v2: " + L.lip_last_error().decode("utf-8", "replace"))
    ws = torch.empty(need, dtype=torch.uint8, device=e.device)
    rc = L.lip_hutchpp_v2(op.ref(), ptr(e), n, int(s1), int(s2), ptr(out), ptr(info), ptr(ws), need, stream())
    op.check(rc, "lip_hutchpp_v2")
    return out[0]


def _cg_matrix(Xfun):
    """Xinv for the *_inv_mvp estimators: CG on every column of a [D,k] block (stochtrace.py:144-147,189-193)."""

    def Xinv(Mx):
        Mx = dev_f32(Mx)
        if Mx.dim() == 1:
            return cg(Xfun, Mx)[0]
        return cg(Xfun, Mx.T.contiguous())[0].T

    return Xinv


def hutchpp_inv_mvp(Xfun, D, seed, num_samples=10, *, eps=None):
    """stochtrace.py:138-148"""
    return hutchpp_mvp(_cg_matrix(Xfun), D, seed, num_samples=num_samples, eps=eps)


def na_hutchpp_dense(X, seed, num_samples=10, *, eps=None):
    """stochtrace.py:151-163"""
    X = dev_f32(X)
    c3 = 0.25
    e = dev_f32(eps) if eps is not None else _rademacher(seed, (num_samples * 4, X.shape[0]))
    ns = e.shape[0] // 4
    S, R, G = e[:ns], e[ns:3 * ns], e[3 * ns:]
    W = X @ S.T
    Zm = X @ R.T
    pin = torch.linalg.pinv(S @ Zm)
    return torch.trace(pin @ (W.T @ Zm)) + (torch.trace(G @ X @ G.T) - torch.trace(G @ Zm @ pin @ W.T @ G.T)) / (c3 * 4 * ns)


def na_hutchpp_mvp(Xfun, D, seed, num_samples=10, dtype=torch.float32, *, eps=None):
    """stochtrace.py:166-180; Xfun maps a matrix."""
    c3 = 0.25
    e = dev_f32(eps) if eps is not None else _rademacher(seed, (num_samples * 4, D))
    ns = e.shape[0] // 4
    S, R, G = e[:ns], e[ns:3 * ns], e[3 * ns:]
    W = dev_f32(Xfun(S.T.contiguous()))
    Zm = dev_f32(Xfun(R.T.contiguous()))
    pin = torch.linalg.pinv(S @ Zm)
    XG = dev_f32(Xfun(G.T.contiguous()))
    return torch.trace(pin @ (W.T @ Zm)) + (torch.trace(G @ XG) - torch.trace(G @ Zm @ pin @ W.T @ G.T)) / (c3 * 4 * ns)


def na_hutchpp_inv_mvp(Xfun, D, seed, num_samples=10, *, eps=None):
    """stochtrace.py:183-194"""
    return na_hutchpp_mvp(_cg_matrix(Xfun), D, seed, num_samples=num_samples, eps=eps)
