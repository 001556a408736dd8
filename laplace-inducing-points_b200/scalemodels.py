"""Architecture descriptors mirroring /root/reference/src/scalemodels.py.

LargeClassifier (scalemodels.py:52-67) runs on the CUDA path.  LeNet5 / ResNet1M (scalemodels.py:11-49,
70-157) are declared so that configs parse, but the conv JVP/VJP kernels are a later SURVEY §8 row:
binding them raises NotImplementedError loudly (no fallback)."""
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
    def apply(self, *a, **k):
        raise NotImplementedError("LeNet5: conv JVP/VJP kernels are not built yet (SURVEY §8a M3)")


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
