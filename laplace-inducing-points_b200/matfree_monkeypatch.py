"""Mirror of /root/reference/src/matfree_monkeypatch.py: the eigenvalue-clipping dense matrix function."""
from __future__ import annotations

from .matfree import DenseFunm, _matfun_name, integrand_funm_sym


def dense_funm_sym_eigh(matfun) -> DenseFunm:
    """matfree_monkeypatch.py:8-22: eigh, clip eigenvalues to >= 1.0 (line 19), V f(L) V^T."""
    return DenseFunm(_matfun_name(matfun), 1.0)


def integrand_funm_sym_logdet(tridiag_sym, /):
    """matfree_monkeypatch.py:25-41: SLQ integrand for logdet built on the clipped eigh."""
    return integrand_funm_sym(DenseFunm("log", 1.0), tridiag_sym)
