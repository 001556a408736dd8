"""lip_b200 — B200-native (sm_100a) drop-in for the matrix-free linearized-Laplace hot path of
nrholm1/Laplace-Inducing-Points.  Module names mirror the reference's `src/` package:

    ggn, lla, stochtrace, sample, matfree_monkeypatch, utils, toymodels, scalemodels, train_inducing (forward objective)
    matfree  (replacements for the third-party matfree / jax.scipy.sparse.linalg.cg routines)

All compute runs in liblip_b200.so (hand-written CUDA behind the C ABI of include/lip_b200.h); importing this
package without the built library, or calling it without a CUDA device, raises — there is no fallback.
"""
from . import _cabi  # noqa: F401

__all__ = ["ggn", "lla", "stochtrace", "sample", "matfree", "matfree_monkeypatch", "utils", "toymodels",
           "scalemodels", "train_inducing", "train_alpha", "evaluate"]


def __getattr__(name):
    if name in __all__:
        import importlib
        return importlib.import_module(f"{__name__}.{name}")
    raise AttributeError(name)
