"""Architecture descriptors mirroring /root/reference/src/scalemodels.py.

LargeClassifier (scalemodels.py:52-67) and LeNet5 (scalemodels.py:11-49) run on the CUDA path.  ResNet1M
(scalemodels.py:70-157) is declared so that configs parse, but its conv/BN/residual JVP/VJP kernels are a later
SURVEY §8 row: binding it raises NotImplementedError loudly (no fallback)."""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Any, Tuple

import numpy as np

from .toymodels import SimpleClassifier, _MLPBase  # noqa: F401  (scalemodels.py:8 re-exports it)

EMPTY_STATS: dict = {}


@dataclass
class LargeClassifier(_MLPBase):
    input_shape: Tuple[int, ...]
    numh: list
    numl: int
    numc: int

    def _dims(self, in_dim):
        return [in_dim] + list(self.numh)[: self.numl] + [self.numc]

    def _in_dim(self, x):
        return int(np.prod(self.input_shape))


class LeNet5:
    """scalemodels.py:11-49.  Runs on the CUDA path as a conv stage program (csrc/lip_cnn.cu)."""
    model_type = "classifier"

    def init(self, seed, x=None):
        """Synthetic initialisation (kernel ~ N(0, 1/fan_in), bias ~ N(0, 0.01^2)); flax-style variables."""
        import math
        rng = np.random.default_rng(int(seed))

        def conv(kh, kw, ci, co):
            return {"bias": (0.01 * rng.standard_normal(co)).astype(np.float32),
                    "kernel": (rng.standard_normal((kh, kw, ci, co)) / math.sqrt(kh * kw * ci)).astype(np.float32)}

        def dense(i, o):
            return {"bias": (0.01 * rng.standard_normal(o)).astype(np.float32),
                    "kernel": (rng.standard_normal((i, o)) / math.sqrt(i)).astype(np.float32)}

        return {"params": {"Conv_0": conv(5, 5, 1, 6), "Conv_1": conv(5, 5, 6, 16), "Dense_0": dense(400, 120),
                           "Dense_1": dense(120, 84), "Dense_2": dense(84, 10)}}

    def apply(self, variables, x, *args, train=False, mutable=False, **kwargs):
        from .ggn import _bind_variables
        return _bind_variables(self, variables, x).outputs()


@dataclass
class ResNet1M:
    num_classes: int = 10

    def apply(self, *a, **k):
        raise NotImplementedError("ResNet1M: conv/BN JVP/VJP kernels are not built yet (SURVEY §8a M4)")


@dataclass
class TrainState:
    """Duck-typed stand-in for flax's TrainState (scalemodels.py:161-163, tests/fixtures.py:65-70)."""
    params: Any
    apply_fn: Any
    batch_stats: Any = field(default_factory=dict)
    alpha: Any = None


def get_model(model_cfg):
    """scalemodels.py:166-185"""
    name = model_cfg["name"]
    if name == "LeNet5":
        return LeNet5()
    if name == "large_classifier":
        return LargeClassifier(tuple(model_cfg["input_shape"]), model_cfg["num_h"], model_cfg["num_l"],
                               model_cfg.get("num_c"))
    if name == "classifier":
        return SimpleClassifier(model_cfg["num_h"], model_cfg["num_l"], model_cfg.get("num_c"))
    if name == "ResNet1":
        return ResNet1M(model_cfg.get("num_c"))
    raise ValueError(f"Unknown model name: {name}")
