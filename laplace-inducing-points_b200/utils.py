"""Mirror of /root/reference/src/utils.py for the hot path: the flat-parameter layout contract."""
from __future__ import annotations

from typing import Callable, Tuple

import numpy as np
import torch

from ._runtime import _require_cuda


def _walk(tree, prefix=()):
    if isinstance(tree, dict):
        for k in sorted(tree.keys()):
            yield from _walk(tree[k], prefix + (k,))
    else:
        yield prefix, tree


def flatten_nn_params(params) -> Tuple[torch.Tensor, Callable]:
    """utils.py:12-17: drop top-level 'logvar'/'batch_stats', then ravel_pytree (sorted-key DFS, row-major leaves).
    Returns (flat fp32 CUDA tensor, unravel_fn)."""
    dev = _require_cuda()
    nn_params = {k: v for k, v in dict(params).items() if k not in ("logvar", "batch_stats")}
    leaves = [(p, l) for p, l in _walk(nn_params)]
    parts, shapes = [], []
    for path, leaf in leaves:
        t = leaf if isinstance(leaf, torch.Tensor) else torch.as_tensor(np.asarray(leaf))
        shapes.append((path, tuple(t.shape)))
        parts.append(t.reshape(-1).to(device=dev, dtype=torch.float32))
    flat = torch.cat(parts) if parts else torch.zeros(0, device=dev)

    def unravel_fn(vec):
        out: dict = {}
        off = 0
        for path, shp in shapes:
            n = int(np.prod(shp)) if len(shp) else 1
            d = out
            for k in path[:-1]:
                d = d.setdefault(k, {})
            d[path[-1]] = vec[off:off + n].reshape(shp)
            off += n
        return out

    return flat, unravel_fn


def count_model_params(params) -> int:
    """utils.py:84"""
    return int(sum(int(np.prod(tuple(l.shape))) if hasattr(l, "shape") else 1 for _, l in _walk(dict(params))))


# ---- minimal optax-protocol optimisers (optax is not in this image; train_inducing.optimize_step / train_alpha.update_alpha
# accept any object with init(params) and update(grads, state, params) -> (updates, new_state); updates are added) ----
class sgd:
    def __init__(self, learning_rate: float):
        self.lr = float(learning_rate)

    def init(self, params):
        return ()

    def update(self, grads, state, params=None):
        return -self.lr * grads, state


class adam:
    """optax.adam(lr, b1=0.9, b2=0.999, eps=1e-8) restated (bias-corrected moments)."""

    def __init__(self, learning_rate: float, b1: float = 0.9, b2: float = 0.999, eps: float = 1e-8):
        self.lr, self.b1, self.b2, self.eps = float(learning_rate), b1, b2, eps

    def init(self, params):
        z = torch.zeros_like(torch.as_tensor(params, dtype=torch.float32))
        return (0, z, z.clone())

    def update(self, grads, state, params=None):
        t, m, v = state
        t = t + 1
        m = self.b1 * m + (1 - self.b1) * grads
        v = self.b2 * v + (1 - self.b2) * grads * grads
        mhat = m / (1 - self.b1 ** t)
        vhat = v / (1 - self.b2 ** t)
        return -self.lr * mhat / (torch.sqrt(vhat) + self.eps), (t, m, v)
