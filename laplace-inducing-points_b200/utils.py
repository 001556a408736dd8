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


# ---- checkpoints (utils.py:20-75; SURVEY §8 row f4: the on-disk formats either side of the hot path) ----------------------------
# Inducing points: plain .npy, `<ckpt_dir>/<name>_<step>.npy` (utils.py:20-43).
# Model states: the reference calls flax.training.checkpoints.save_checkpoint(target=TrainState, prefix=prefix + "_") (utils.py:46-73),
# i.e. ONE file `<ckpt_dir>/<prefix>_<step>` holding flax.serialization.to_bytes(state) = msgpack of the state dict
# {"step", "params", "opt_state", "batch_stats", ...} with every array as msgpack ExtType(1, packb((shape, dtype_name, raw_bytes)))
# and numpy scalars as ExtType(3, packb((dtype_name, raw_bytes))).  flax is not in this image and the reference ships no checkpoint
# (checkpoint/ is git-ignored), so this layout is restated from flax's published serialization format — PARITY UNPINNED — and is
# exercised by round trips plus a hand-assembled byte string in tests/test_checkpoints.py.
_EXT_NDARRAY, _EXT_NPSCALAR = 1, 3


def save_array_checkpoint(array, ckpt_dir, name, step):
    """utils.py:20-29"""
    import os
    ckpt_dir = os.path.abspath(ckpt_dir)
    os.makedirs(ckpt_dir, exist_ok=True)
    filename = os.path.join(ckpt_dir, f"{name}_{step}.npy")
    np.save(filename, np.asarray(array.detach().cpu() if isinstance(array, torch.Tensor) else array))
    return filename


def load_array_checkpoint(ckpt_dir, name, step, device=None):
    """utils.py:32-43: returns the array on the current CUDA device (jax.device_put in the reference), or as numpy with device='cpu'."""
    import os
    filename = os.path.join(os.path.abspath(ckpt_dir), f"{name}_{step}.npy")
    if not os.path.exists(filename):
        raise FileNotFoundError(f"Checkpoint file {filename} not found")
    array = np.load(filename)
    if device == "cpu":
        return array
    from ._runtime import dev_f32
    return dev_f32(array) if array.dtype.kind == "f" else torch.as_tensor(array).to(dev_f32(np.zeros(1)).device)


def _ext_hook(code, data):
    import msgpack
    if code == _EXT_NDARRAY:
        shape, dtype_name, buf = msgpack.unpackb(data, raw=False, strict_map_key=False)
        return np.frombuffer(buf, dtype=np.dtype(dtype_name)).reshape(tuple(shape)).copy()
    if code == _EXT_NPSCALAR:
        dtype_name, buf = msgpack.unpackb(data, raw=False, strict_map_key=False)
        return np.frombuffer(buf, dtype=np.dtype(dtype_name))[0]
    raise ValueError(f"unsupported msgpack extension type {code} in checkpoint")


def _ext_default(obj):
    import msgpack
    if isinstance(obj, torch.Tensor):
        obj = obj.detach().cpu().numpy()
    if isinstance(obj, np.ndarray):
        a = np.ascontiguousarray(obj)
        return msgpack.ExtType(_EXT_NDARRAY, msgpack.packb((list(a.shape), a.dtype.name, a.tobytes()), use_bin_type=True))
    if isinstance(obj, np.generic):
        return msgpack.ExtType(_EXT_NPSCALAR, msgpack.packb((obj.dtype.name, obj.tobytes()), use_bin_type=True))
    raise TypeError(f"cannot serialise {type(obj)}")


def state_dict_to_bytes(state_dict) -> bytes:
    """flax.serialization.msgpack_serialize of a nested dict of arrays."""
    import msgpack
    return msgpack.packb(state_dict, default=_ext_default, use_bin_type=True, strict_types=True)


def state_dict_from_bytes(data: bytes):
    """flax.serialization.msgpack_restore: nested dict with numpy leaves."""
    import msgpack
    return msgpack.unpackb(data, ext_hook=_ext_hook, raw=False, strict_map_key=False)


def _latest(ckpt_dir, prefix):
    import os, re
    best, best_step = None, None
    for fn in os.listdir(ckpt_dir):
        m = re.fullmatch(re.escape(prefix) + r"(\d+(?:\.\d+)?)", fn)
        if m and (best_step is None or float(m.group(1)) > best_step):
            best, best_step = fn, float(m.group(1))
    return None if best is None else os.path.join(ckpt_dir, best)


def save_checkpoint(train_state, ckpt_dir, prefix, step):
    """utils.py:46-60: `<ckpt_dir>/<prefix>_<step>`; older steps of the same prefix are removed (flax keep=1, overwrite=True).
    The reference serialises the WHOLE flax TrainState (step, params, opt_state, batch_stats) and restores it with
    target=TrainState, which raises on a missing field: `opt_state` is therefore written whenever the state carries one (any
    nested dict / list / tuple of arrays; tuples and lists become flax's {'0': ..., '1': ...} index dicts).  A state without an
    optimiser (the evaluation-side TrainState of this package) writes an empty opt_state, which flax restores only into a target
    whose opt_state is empty as well — such files are for this package's load_checkpoint and for target=None readers."""
    import os
    ckpt_dir = os.path.abspath(ckpt_dir)
    os.makedirs(ckpt_dir, exist_ok=True)
    sd = {"step": int(step), "params": _to_numpy_tree(train_state.params),
          "opt_state": _to_numpy_tree(getattr(train_state, "opt_state", None) or {}),
          "batch_stats": _to_numpy_tree(getattr(train_state, "batch_stats", {}) or {})}
    old = _latest(ckpt_dir, prefix + "_")
    path = os.path.join(ckpt_dir, f"{prefix}_{step}")
    with open(path, "wb") as f:
        f.write(state_dict_to_bytes(sd))
    if old and os.path.abspath(old) != path:
        os.remove(old)
    return path


def load_checkpoint(ckpt_dir, prefix, target=None):
    """utils.py:63-75: restores the LATEST step of `<prefix>_*`.  With a target state, returns a copy of it whose params / batch_stats
    come from the file (shapes checked leaf by leaf); without one, the raw state dict (flax's behaviour for target=None)."""
    import dataclasses
    import os
    path = _latest(os.path.abspath(ckpt_dir), prefix + "_")
    if path is None:
        return target                                    # flax returns the target unchanged when nothing is found
    with open(path, "rb") as f:
        sd = state_dict_from_bytes(f.read())
    if target is None:
        return sd
    params = _restore_like(target.params, sd["params"], "params")
    bs = _restore_like(getattr(target, "batch_stats", {}) or {}, sd.get("batch_stats", {}) or {}, "batch_stats")
    fields = {"params": params, "batch_stats": bs}
    if getattr(target, "opt_state", None) and sd.get("opt_state"):
        fields["opt_state"] = _restore_like(_to_numpy_tree(target.opt_state), sd["opt_state"], "opt_state")
    return dataclasses.replace(target, **{k: v for k, v in fields.items() if hasattr(target, k)})


def _to_numpy_tree(t):
    if isinstance(t, dict):
        return {str(k): _to_numpy_tree(v) for k, v in t.items()}
    if isinstance(t, (list, tuple)):                  # flax.serialization: sequences are dicts keyed by their index
        if hasattr(t, "_asdict"):
            return {str(k): _to_numpy_tree(v) for k, v in t._asdict().items()}
        return {str(i): _to_numpy_tree(v) for i, v in enumerate(t)}
    if t is None:
        return {}
    if isinstance(t, torch.Tensor):
        return t.detach().cpu().numpy()
    return np.asarray(t)


def _restore_like(like, loaded, where):
    if isinstance(like, dict):
        if not isinstance(loaded, dict) or sorted(map(str, like)) != sorted(map(str, loaded)):
            raise ValueError(f"checkpoint {where}: keys {sorted(map(str, loaded)) if isinstance(loaded, dict) else type(loaded)} "
                             f"do not match the target's {sorted(map(str, like))}")
        return {k: _restore_like(v, loaded[str(k)], f"{where}/{k}") for k, v in like.items()}
    arr = np.asarray(loaded)
    if arr.size == 1 and int(np.size(like)) == 1:
        arr = arr.reshape(np.shape(like))              # 0-d leaves (optimizer step counts) come back as one-element arrays
    if tuple(arr.shape) != tuple(np.shape(like)):
        raise ValueError(f"checkpoint {where}: shape {tuple(arr.shape)} does not match the target's {tuple(np.shape(like))}")
    return arr
