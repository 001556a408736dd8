"""Oracle model zoo: torch float64 CPU restatements of the reference's flax modules.

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

Reference being restated (forward math only; flax itself is absent from this image):
  * SimpleRegressor   /root/reference/src/toymodels.py:4-24   numl x [Dense -> gelu(tanh approx)] -> Dense(1)
  * SimpleClassifier  /root/reference/src/toymodels.py:27-37  numl x [Dense -> tanh] -> Dense(numc)
  * LargeClassifier   /root/reference/src/scalemodels.py:52-67 flatten -> [Dense -> tanh]* -> Dense(numc)
  * LeNet5            /root/reference/src/scalemodels.py:11-49
  * ResNet1M          /root/reference/src/scalemodels.py:70-157 (BatchNorm in eval mode, ggn.py:52)

Parameter layout = /root/reference/src/utils.py:12-17 (flatten_nn_params): drop the top-level
'logvar' / 'batch_stats' entries, then jax.flatten_util.ravel_pytree == depth-first traversal with
dict keys sorted lexicographically at every level, each leaf raveled row-major.
Flax leaf names: Dense {bias[out], kernel[in,out]}, Conv {bias[cout]?, kernel[kh,kw,cin,cout]},
BatchNorm {bias[c], scale[c]} (+ batch_stats {mean[c], var[c]}).
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field
from typing import Dict, List, Tuple

import numpy as np
import torch
import torch.nn.functional as F

DT = torch.float64


# --------------------------------------------------------------------------------------
# pytree flatten (ravel_pytree restatement)
# --------------------------------------------------------------------------------------
def _walk(tree, prefix=()):
    """Depth-first, sorted-key traversal yielding (path, leaf)."""
    if isinstance(tree, dict):
        for k in sorted(tree.keys()):
            yield from _walk(tree[k], prefix + (k,))
    else:
        yield prefix, tree


def flatten_nn_params(params: dict):
    """utils.py:12-17.  Returns (flat float64 numpy vector, unravel_fn)."""
    nn_params = {k: v for k, v in params.items() if k not in ("logvar", "batch_stats")}
    leaves = [(p, np.asarray(l)) for p, l in _walk(nn_params)]
    flat = np.concatenate([l.reshape(-1).astype(np.float64) for _, l in leaves]) if leaves else np.zeros(0)
    shapes = [(p, l.shape) for p, l in leaves]

    def unravel(vec):
        out: dict = {}
        off = 0
        for path, shp in shapes:
            n = int(np.prod(shp)) if len(shp) else 1
            leaf = vec[off:off + n].reshape(shp)
            off += n
            d = out
            for k in path[:-1]:
                d = d.setdefault(k, {})
            d[path[-1]] = leaf
        return out

    return flat, unravel


def _strip_params_level(tree: dict) -> dict:
    """Toy layout wraps the module dict in one extra {'params': ...} level (main.py:191-194)."""
    keys = [k for k in tree.keys() if k not in ("logvar", "batch_stats")]
    if keys == ["params"]:
        return tree["params"]
    return {k: tree[k] for k in keys}


# --------------------------------------------------------------------------------------
# activations (flax defaults)
# --------------------------------------------------------------------------------------
def gelu_tanh(x):
    # flax.linen.gelu(approximate=True) == jax.nn.gelu default
    c = math.sqrt(2.0 / math.pi)
    return 0.5 * x * (1.0 + torch.tanh(c * (x + 0.044715 * x ** 3)))


_ACT = {"tanh": torch.tanh, "gelu": gelu_tanh, "relu": torch.relu}


# --------------------------------------------------------------------------------------
# model descriptions
# --------------------------------------------------------------------------------------
@dataclass
class OracleModel:
    """kind in {'regressor_mlp','classifier_mlp','large_classifier','lenet5','resnet1m'}"""
    kind: str
    in_shape: Tuple[int, ...]
    hidden: List[int] = field(default_factory=list)
    num_out: int = 1
    model_type: str = "classifier"  # 'regressor' | 'classifier'

    # ---- parameter initialisation (synthetic, SURVEY 8d: kernel~N(0,1/fan_in), bias~N(0,0.01^2)) ----
    def init(self, seed: int) -> dict:
        rng = np.random.default_rng(seed)

        def dense(i, o):
            return {"bias": 0.01 * rng.standard_normal(o), "kernel": rng.standard_normal((i, o)) / math.sqrt(i)}

        def conv(kh, kw, ci, co, bias=True):
            d = {"kernel": rng.standard_normal((kh, kw, ci, co)) / math.sqrt(kh * kw * ci)}
            if bias:
                d["bias"] = 0.01 * rng.standard_normal(co)
            return d

        def bn(c):
            return ({"bias": 0.1 * rng.standard_normal(c), "scale": 1.0 + 0.1 * rng.standard_normal(c)},
                    {"mean": 0.1 * rng.standard_normal(c), "var": 1.0 + 0.1 * np.abs(rng.standard_normal(c))})

        p: dict = {}
        stats: dict = {}
        if self.kind == "linear1d":
            p = {"W": 0.1 * rng.standard_normal(()), "b": 0.1 * rng.standard_normal(())}
        elif self.kind in ("regressor_mlp", "classifier_mlp", "large_classifier"):
            dims = [int(np.prod(self.in_shape))] + list(self.hidden) + [self.num_out]
            for j in range(len(dims) - 1):
                p[f"Dense_{j}"] = dense(dims[j], dims[j + 1])
        elif self.kind == "lenet5":
            p["Conv_0"] = conv(5, 5, 1, 6)
            p["Conv_1"] = conv(5, 5, 6, 16)
            p["Dense_0"] = dense(400, 120)
            p["Dense_1"] = dense(120, 84)
            p["Dense_2"] = dense(84, 10)
        elif self.kind == "resnet1m":
            cin = 3
            p["Conv_0"] = conv(3, 3, cin, 32, bias=False)
            p["BatchNorm_0"], stats["BatchNorm_0"] = bn(32)
            chans = [(32, 1), (32, 1), (32, 1), (64, 2), (64, 1), (64, 1), (128, 2), (128, 1), (128, 1)]
            c_prev = 32
            for bi, (c, s) in enumerate(chans):
                blk: dict = {}
                bst: dict = {}
                blk["Conv_0"] = conv(3, 3, c_prev, c, bias=False)
                blk["BatchNorm_0"], bst["BatchNorm_0"] = bn(c)
                blk["Conv_1"] = conv(3, 3, c, c, bias=False)
                blk["BatchNorm_1"], bst["BatchNorm_1"] = bn(c)
                if s != 1 or c_prev != c:
                    blk["Conv_2"] = conv(1, 1, c_prev, c, bias=False)
                    blk["BatchNorm_2"], bst["BatchNorm_2"] = bn(c)
                p[f"BasicBlock_{bi}"] = blk
                stats[f"BasicBlock_{bi}"] = bst
                c_prev = c
            p["Dense_0"] = dense(128, self.num_out)
        else:
            raise ValueError(self.kind)
        p = _to_f32_tree(p)
        stats = _to_f32_tree(stats)
        return {"params": p, "batch_stats": stats}

    # ---- forward: theta is the flat float64 torch vector; x is [n, *in_shape] ----
    def forward(self, tree: dict, x: torch.Tensor, batch_stats: dict | None = None) -> torch.Tensor:
        """tree: unravelled (possibly toy-wrapped) dict of torch tensors."""
        p = _strip_params_level(tree)
        if self.kind == "linear1d":  # tests/fixtures.py:41-48  mu = W*x + b (scalars)
            return p["W"] * x + p["b"]
        if self.kind in ("regressor_mlp", "classifier_mlp", "large_classifier"):
            act = _ACT["gelu"] if self.kind == "regressor_mlp" else _ACT["tanh"]
            h = x.reshape(x.shape[0], -1)
            n = len(p)
            for j in range(n):
                d = p[f"Dense_{j}"]
                h = h @ d["kernel"] + d["bias"]
                if j < n - 1:
                    h = act(h)
            return h
        if self.kind == "lenet5":
            h = x.permute(0, 3, 1, 2)  # NHWC -> NCHW
            h = F.pad(h, (2, 2, 2, 2))
            h = _conv(h, p["Conv_0"], 1, (0, 0, 0, 0))
            h = F.avg_pool2d(torch.relu(h), 2, 2)
            h = _conv(h, p["Conv_1"], 1, (0, 0, 0, 0))
            h = F.avg_pool2d(torch.relu(h), 2, 2)
            h = h.permute(0, 2, 3, 1).reshape(h.shape[0], -1)  # flatten in HWC order
            h = torch.relu(h @ p["Dense_0"]["kernel"] + p["Dense_0"]["bias"])
            h = torch.relu(h @ p["Dense_1"]["kernel"] + p["Dense_1"]["bias"])
            return h @ p["Dense_2"]["kernel"] + p["Dense_2"]["bias"]
        if self.kind == "resnet1m":
            bs = batch_stats
            h = x.permute(0, 3, 1, 2)
            if h.shape[1] == 1:
                h = h.repeat(1, 3, 1, 1)
            h = _conv_same(h, p["Conv_0"], 1)
            h = torch.relu(_bn(h, p["BatchNorm_0"], bs["BatchNorm_0"]))
            for bi in range(9):
                blk, bst = p[f"BasicBlock_{bi}"], bs[f"BasicBlock_{bi}"]
                stride = 2 if "Conv_2" in blk else 1
                r = h
                y = _conv_same(h, blk["Conv_0"], stride)
                y = torch.relu(_bn(y, blk["BatchNorm_0"], bst["BatchNorm_0"]))
                y = _conv_same(y, blk["Conv_1"], 1)
                y = _bn(y, blk["BatchNorm_1"], bst["BatchNorm_1"])
                if "Conv_2" in blk:
                    r = _conv_same(r, blk["Conv_2"], stride)
                    r = _bn(r, blk["BatchNorm_2"], bst["BatchNorm_2"])
                h = torch.relu(y + r)
            h = h.mean(dim=(2, 3))
            return h @ p["Dense_0"]["kernel"] + p["Dense_0"]["bias"]
        raise ValueError(self.kind)


def _to_f32_tree(t):
    if isinstance(t, dict):
        return {k: _to_f32_tree(v) for k, v in t.items()}
    return np.asarray(t, dtype=np.float32)


def _conv(h, cp, stride, pad):
    w = cp["kernel"].permute(3, 2, 0, 1)  # HWIO -> OIHW (cross-correlation, as lax.conv)
    h = F.pad(h, pad)
    return F.conv2d(h, w, cp.get("bias"), stride=stride)


def _same_pad(n, k, s):
    out = -(-n // s)
    total = max((out - 1) * s + k - n, 0)
    return total // 2, total - total // 2


def _conv_same(h, cp, stride):
    kh, kw = cp["kernel"].shape[0], cp["kernel"].shape[1]
    pt, pb = _same_pad(h.shape[2], kh, stride)
    pl, pr = _same_pad(h.shape[3], kw, stride)
    return _conv(h, cp, stride, (pl, pr, pt, pb))


def _bn(h, bp, st, eps=1e-5):
    mul = torch.rsqrt(st["var"] + eps) * bp["scale"]
    return (h - st["mean"][None, :, None, None]) * mul[None, :, None, None] + bp["bias"][None, :, None, None]


# --------------------------------------------------------------------------------------
# constructors named after the reference's module classes
# --------------------------------------------------------------------------------------
def Linear1D() -> OracleModel:
    """the hand-rolled 1-layer linear 'model' of tests/fixtures.py:29-70"""
    return OracleModel("linear1d", (1,), [], 1, "regressor")


def SimpleRegressor(numh: int, numl: int, in_dim: int = 1) -> OracleModel:
    return OracleModel("regressor_mlp", (in_dim,), [numh] * numl, 1, "regressor")


def SimpleClassifier(numh: int, numl: int, numc: int, in_dim: int = 2) -> OracleModel:
    return OracleModel("classifier_mlp", (in_dim,), [numh] * numl, numc, "classifier")


def LargeClassifier(input_shape, numh, numl, numc) -> OracleModel:
    return OracleModel("large_classifier", tuple(input_shape), list(numh)[:numl], numc, "classifier")


def LeNet5() -> OracleModel:
    return OracleModel("lenet5", (28, 28, 1), [], 10, "classifier")


def ResNet1M(num_classes: int = 10, in_shape=(32, 32, 3)) -> OracleModel:
    return OracleModel("resnet1m", tuple(in_shape), [], num_classes, "classifier")


# --------------------------------------------------------------------------------------
# a reference-shaped "state" (duck-typed like tests/fixtures.py:65-70)
# --------------------------------------------------------------------------------------
class OracleState:
    """Holds .params (nested dict of numpy arrays, reference layout), .batch_stats, .model, .logvar."""

    def __init__(self, model: OracleModel, variables: dict, toy_layout: bool = False, logvar: float = 0.0):
        self.model = model
        self.batch_stats = variables.get("batch_stats", {})
        if model.model_type == "regressor":
            # main.py:191-194 / fixtures.py:56-63: params == {'params': {...}, 'logvar': {'logvar': s}}
            self.params = {"params": variables["params"], "logvar": {"logvar": np.float32(logvar)}}
        elif toy_layout:
            self.params = {"params": variables["params"]}
        else:
            self.params = variables["params"]  # scale layout (scale_experiments/train.py:86)

    @property
    def logvar(self) -> float:
        return float(self.params["logvar"]["logvar"]) if "logvar" in self.params else 0.0

    def flat(self):
        return flatten_nn_params(self.params)

    def f_theta(self, theta: torch.Tensor, x: torch.Tensor) -> torch.Tensor:
        """Model output [n, K] as a differentiable function of the flat float64 parameter vector."""
        tree = _unravel_torch(theta, self.params)
        bs = _tree_to_torch(self.batch_stats)
        return self.model.forward(tree, x, bs)


def _tree_to_torch(t):
    if isinstance(t, dict):
        return {k: _tree_to_torch(v) for k, v in t.items()}
    return torch.as_tensor(np.asarray(t), dtype=DT)


def _unravel_torch(theta: torch.Tensor, params: dict):
    nn_params = {k: v for k, v in params.items() if k not in ("logvar", "batch_stats")}
    out: dict = {}
    off = 0
    for path, leaf in _walk(nn_params):
        shp = np.asarray(leaf).shape
        n = int(np.prod(shp)) if len(shp) else 1
        d = out
        for k in path[:-1]:
            d = d.setdefault(k, {})
        d[path[-1]] = theta[off:off + n].reshape(shp)
        off += n
    return out
