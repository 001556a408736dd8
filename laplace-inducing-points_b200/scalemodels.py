"""Architecture descriptors mirroring /root/reference/src/scalemodels.py.

LargeClassifier (scalemodels.py:52-67), LeNet5 (scalemodels.py:11-49) and ResNet1M (scalemodels.py:70-157, eval-mode
BatchNorm) run on the CUDA path; anything else raises loudly (no fallback)."""
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
    """scalemodels.py:112-157 (+ BasicBlock :70-109).  Runs on the CUDA path as a residual conv program
    (csrc/lip_resnet.cu), BatchNorm in eval mode with the running statistics of `variables['batch_stats']`."""
    num_classes: int = 10
    model_type = "classifier"

    def apply(self, variables, x, *args, train=False, mutable=False, **kwargs):
        if train:
            raise NotImplementedError("ResNet1M: only eval-mode BatchNorm (use_running_average) is on the hot path (ggn.py:52)")
        from ._runtime import BoundModel, ResNetProgramSpec, dev_f32
        from .ggn import _strip
        from .utils import flatten_nn_params
        xt = dev_f32(x)
        if xt.dim() == 3:
            xt = xt[None]
        if xt.shape[-1] == 1:
            xt = xt.repeat(1, 1, 1, 3)
        params = {k: v for k, v in dict(variables).items() if k != "batch_stats"}
        spec = ResNetProgramSpec(_strip(params), dict(variables).get("batch_stats"), tuple(xt.shape[1:]), "classifier")
        theta, _ = flatten_nn_params(params)
        return BoundModel(spec, theta, xt, 0.0).outputs()


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
