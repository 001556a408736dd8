"""Architecture descriptors mirroring /root/reference/src/toymodels.py (flax modules SimpleRegressor,
SimpleClassifier).  `apply` runs the forward pass through the CUDA library (lip_model_bind caches it)."""
from __future__ import annotations

import math
from dataclasses import dataclass

import numpy as np
import torch


def _dense_init(rng, nin, nout):
    return {"bias": (0.01 * rng.standard_normal(nout)).astype(np.float32),
            "kernel": (rng.standard_normal((nin, nout)) / math.sqrt(nin)).astype(np.float32)}


class _MLPBase:
    activation = "tanh"
    model_type = "classifier"

    def _dims(self, in_dim):
        raise NotImplementedError

    def init(self, seed, x):
        """Synthetic initialisation (kernel ~ N(0, 1/fan_in), bias ~ N(0, 0.01^2)); returns flax-style variables."""
        rng = np.random.default_rng(int(seed))
        x = np.asarray(x.detach().cpu() if isinstance(x, torch.Tensor) else x)
        in_dim = self._in_dim(x)
        dims = self._dims(in_dim)
        params = {f"Dense_{j}": _dense_init(rng, dims[j], dims[j + 1]) for j in range(len(dims) - 1)}
        out = {"params": params}
        if self.model_type == "regressor":
            out["logvar"] = {"logvar": np.float32(0.0)}
        return out

    def _in_dim(self, x):
        return int(x.shape[-1])

    def apply(self, variables, x, *args, return_logvar=False, train=False, mutable=False, **kwargs):
        from .ggn import _bind_variables  # local import: avoids a cycle
        bm = _bind_variables(self, variables, x)
        out = bm.outputs()
        if self.model_type == "regressor" and return_logvar:
            return out, variables["logvar"]["logvar"]
        return out


@dataclass
class SimpleRegressor(_MLPBase):
    """toymodels.py:4-24: numl x [Dense(numh) -> gelu(tanh approx)] -> Dense(1); logvar in its own collection."""
    numh: int
    numl: int
    activation = "gelu"
    model_type = "regressor"

    def _dims(self, in_dim):
        return [in_dim] + [self.numh] * self.numl + [1]


@dataclass
class SimpleClassifier(_MLPBase):
    """toymodels.py:27-37: numl x [Dense(numh) -> tanh] -> Dense(numc)."""
    numh: int
    numl: int
    numc: int

    def _dims(self, in_dim):
        return [in_dim] + [self.numh] * self.numl + [self.numc]
