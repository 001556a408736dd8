"""Host-side plumbing over the C ABI: device tensors (torch), workspaces, the bound-model handle."""
from __future__ import annotations

import ctypes as C
import math
import weakref
from typing import List, Optional, Sequence, Tuple

import torch

from . import _cabi as cabi


def _require_cuda() -> torch.device:
    if not torch.cuda.is_available():
        raise cabi.LipError("lip_b200 needs a CUDA device (sm_100a); there is no CPU fallback for this path")
    return torch.device("cuda", torch.cuda.current_device())


def dev_f32(x, device=None) -> torch.Tensor:
    device = device or _require_cuda()
    if isinstance(x, torch.Tensor):
        if x.device != device:   # move first (possibly a narrow dtype from pinned host memory), cast on the device
            x = x.to(device=device, non_blocking=True)
        return x.to(dtype=torch.float32).contiguous()
    import numpy as np
    return torch.as_tensor(np.asarray(x), dtype=torch.float32).to(device).contiguous()


def pad4(n: int) -> int:
    return (int(n) + 3) // 4 * 4


def exact_tf32_block(rows: int, n: int, device=None) -> torch.Tensor:
    """An uninitialised [rows, n] float32 view of a [rows, pad4(n)] buffer, marked as holding exactly-TF32 data: the layout and the
    promise lip_ggn_vp_ex needs to read a probe block in place (LIP_PROBES_EXACT_TF32).  Only code that then FILLS it with such
    data (+-1 Rademacher probes, one-hot rows) may use it: stochtrace._rademacher, unpack_rademacher, sampler_rademacher."""
    device = device or _require_cuda()
    buf = torch.empty(rows, pad4(n), device=device, dtype=torch.float32)
    view = buf[:, :n]
    view._lip_exact_tf32 = True
    return view


def is_exact_tf32(t) -> bool:
    return isinstance(t, torch.Tensor) and getattr(t, "_lip_exact_tf32", False)


def row_strided_f32(x, device=None) -> torch.Tensor:
    """dev_f32 that keeps a 2-D CUDA float32 tensor with unit column stride as it is (no .contiguous() copy of padded rows)."""
    if isinstance(x, torch.Tensor) and x.is_cuda and x.dtype == torch.float32 and x.dim() == 2 and x.stride(1) == 1 \
            and x.stride(0) >= x.shape[1] and (device is None or x.device == device):
        return x
    return dev_f32(x, device)


def ptr(t: Optional[torch.Tensor]):
    return C.c_void_p(0 if t is None else t.data_ptr())


def stream() -> C.c_void_p:
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


class Scratch:
    """A growable byte buffer (caller-owned scratch of the C ABI)."""

    def __init__(self):
        self.buf: Optional[torch.Tensor] = None

    def get(self, nbytes: int) -> Tuple[C.c_void_p, int]:
        nbytes = int(nbytes)
        if self.buf is None or self.buf.numel() < nbytes:
            self.buf = torch.empty(max(nbytes, 1 << 20), dtype=torch.uint8, device=_require_cuda())
        return C.c_void_p(self.buf.data_ptr()), self.buf.numel()


_GLOBAL_SCRATCH = Scratch()


def scratch(nbytes: int):
    return _GLOBAL_SCRATCH.get(nbytes)


class _SharedWorkspace:
    """Operator workspace shared by every BoundModel that issues work on the same (device, stream): calls on one stream are
    ordered, so they can reuse one buffer — an objective that binds three models (S_X, W_z, W_x) needs the LARGEST of their
    workspaces, not the sum (a ResNet1M at M = 100 asks for tens of GB per 256-probe block).  Different streams get different buffers."""

    def __init__(self):
        self.bufs = {}

    def get(self, nbytes: int) -> Tuple[C.c_void_p, int]:
        dev = torch.cuda.current_device()
        key = (dev, torch.cuda.current_stream().cuda_stream)
        buf = self.bufs.get(key)
        if buf is None or buf.numel() < int(nbytes):
            self.bufs.pop(key, None)
            del buf
            buf = torch.empty(max(int(nbytes), 1 << 20), dtype=torch.uint8, device=torch.device("cuda", dev))
            self.bufs[key] = buf
            if len(self.bufs) > 8:                       # streams come and go (CUDA-graph capture, pipelines): keep the table small
                self.bufs.pop(next(iter(self.bufs)))
        return C.c_void_p(buf.data_ptr()), buf.numel()


_SHARED_WS = _SharedWorkspace()


class MLPSpec:
    """Architecture of an MLP extracted from a reference-layout parameter tree (Dense_i: bias, kernel)."""

    def __init__(self, dims: Sequence[int], act: int, model_type: str):
        self.dims = list(dims)
        self.act = act
        self.model_type = model_type
        off = 0
        self.layers = []
        for i in range(len(self.dims) - 1):
            nin, nout = self.dims[i], self.dims[i + 1]
            self.layers.append((nin, nout, off, off + nout))  # bias then kernel (sorted keys)
            off += nout + nin * nout
        self.num_params = off

    def descs(self):
        out = []
        n = len(self.layers)
        for i, (nin, nout, boff, woff) in enumerate(self.layers):
            out.append(cabi.LayerDesc(cabi.OP_DENSE, nin, nout, boff, woff))
            if i < n - 1:
                out.append(cabi.LayerDesc(self.act, 0, 0, 0, 0))
        arr = (cabi.LayerDesc * len(out))(*out)
        return arr, len(out)


class ConvProgramSpec:
    """A conv stage program (INPUT, ZEROPAD, CONV2D, act, AVGPOOL2, FLATTEN, DENSE ...) with parameter offsets taken
    from the reference's flat order (sorted module names: Conv_i before Dense_i; inside a module bias before kernel).

    ops: list of tuples
        ("input", H, W, C) | ("pad", p) | ("conv", module_name, kh, kw, cin, cout) | ("act", op) | ("pool",) |
        ("flatten",) | ("dense", module_name, nin, nout)
    """

    def __init__(self, ops, model_type: str, name: str = "conv"):
        self.ops = list(ops)
        self.model_type = model_type
        self.name = name
        sizes = {}
        for op in self.ops:
            if op[0] == "conv":
                _, mod, kh, kw, cin, cout = op
                sizes[mod] = (cout, kh * kw * cin * cout)
            elif op[0] == "dense":
                _, mod, nin, nout = op
                sizes[mod] = (nout, nin * nout)
        self.offsets = {}
        off = 0
        for mod in sorted(sizes):                       # ravel_pytree: sorted keys; leaves 'bias' < 'kernel'
            nb, nk = sizes[mod]
            self.offsets[mod] = (off, off + nb)
            off += nb + nk
        self.num_params = off
        inp = self.ops[0]
        if inp[0] != "input":
            raise ValueError("a conv program starts with ('input', H, W, C)")
        self.in_shape = tuple(inp[1:4])
        self.in_features = int(inp[1] * inp[2] * inp[3])
        last = [op for op in self.ops if op[0] == "dense"][-1]
        self.num_outputs = int(last[3])
        self.layers = [op for op in self.ops if op[0] in ("conv", "dense")]

    def descs(self):
        out = []
        for op in self.ops:
            k = op[0]
            if k == "input":
                out.append(cabi.LayerDesc(cabi.OP_INPUT, op[3], 0, 0, 0, op[1], op[2], 0, 0))
            elif k == "pad":
                out.append(cabi.LayerDesc(cabi.OP_ZEROPAD, 0, 0, 0, 0, 0, 0, 0, op[1]))
            elif k == "conv":
                _, mod, kh, kw, cin, cout = op
                boff, woff = self.offsets[mod]
                out.append(cabi.LayerDesc(cabi.OP_CONV2D, cin, cout, boff, woff, kh, kw, 1, 0))
            elif k == "act":
                out.append(cabi.LayerDesc(op[1], 0, 0, 0, 0))
            elif k == "pool":
                out.append(cabi.LayerDesc(cabi.OP_AVGPOOL2, 0, 0, 0, 0))
            elif k == "flatten":
                out.append(cabi.LayerDesc(cabi.OP_FLATTEN, 0, 0, 0, 0))
            elif k == "dense":
                _, mod, nin, nout = op
                boff, woff = self.offsets[mod]
                out.append(cabi.LayerDesc(cabi.OP_DENSE, nin, nout, boff, woff))
            else:
                raise ValueError(f"unknown program op {op!r}")
        arr = (cabi.LayerDesc * len(out))(*out)
        return arr, len(out)


def _tree_walk(tree, prefix=()):
    if isinstance(tree, dict):
        for k in sorted(tree.keys()):
            yield from _tree_walk(tree[k], prefix + (k,))
    else:
        yield prefix, tree


class ResNetProgramSpec:
    """ResNet1M (scalemodels.py:70-157) as a residual conv program.  Geometry and parameter offsets are read off the
    parameter tree itself (flat order = sorted-key DFS, utils.py:12-17): stem Conv_0 + BatchNorm_0, BasicBlock_i with
    Conv_0/1 (+ Conv_2 = 1x1 shortcut, which is also how the stride-2 blocks are recognised, scalemodels.py:101-109),
    Dense_0.  BatchNorm running statistics (state.batch_stats) are constants of the program."""

    def __init__(self, tree, batch_stats, in_shape, model_type: str):
        import numpy as np
        self.model_type = model_type
        self.name = "ResNet1M"
        self.in_shape = tuple(int(x) for x in in_shape)
        if len(self.in_shape) != 3:
            raise ValueError(f"ResNet1M expects NHWC images, got per-point shape {self.in_shape}")
        self.in_features = int(np.prod(self.in_shape))
        offs, off = {}, 0
        for path, leaf in _tree_walk(tree):
            offs[path] = off
            off += int(np.prod(tuple(leaf.shape))) if len(tuple(leaf.shape)) else 1
        self.num_params = off
        stats_tree = dict(batch_stats or {})
        if list(stats_tree.keys()) == ["batch_stats"]:
            stats_tree = dict(stats_tree["batch_stats"])
        H, W, C = self.in_shape
        self.ops = [cabi.LayerDesc(cabi.OP_INPUT, C, 0, 0, 0, H, W, 0, 0)]
        self.layers = []
        stats = []

        def sub(t, path):
            for k in path:
                if not isinstance(t, dict) or k not in t:
                    raise ValueError(f"ResNet1M: missing {'/'.join(path)} in the parameter / batch_stats tree")
                t = t[k]
            return t

        def conv_bn(path, conv, bn, cin, hin, stride, op_conv, op_bn):
            kern = sub(tree, path + (conv, "kernel"))
            if "bias" in sub(tree, path + (conv,)):
                raise ValueError(f"ResNet1M: {'/'.join(path + (conv,))} has a bias (use_bias=False expected)")
            kh, kw, ci, co = (int(x) for x in kern.shape)
            if ci != cin or kh != kw:
                raise ValueError(f"ResNet1M: {'/'.join(path + (conv,))} kernel {tuple(kern.shape)} does not fit {cin} input channels")
            hout = -(-hin // stride)
            pad = max((hout - 1) * stride + kh - hin, 0) // 2            # XLA 'SAME': the smaller half goes first
            self.ops.append(cabi.LayerDesc(op_conv, ci, co, -1, offs[path + (conv, "kernel")], kh, kw, stride, pad))
            bnp = sub(tree, path + (bn,))
            if tuple(bnp["scale"].shape) != (co,) or tuple(bnp["bias"].shape) != (co,):
                raise ValueError(f"ResNet1M: {'/'.join(path + (bn,))} does not have {co} channels")
            self.ops.append(cabi.LayerDesc(op_bn, co, co, offs[path + (bn, "bias")], offs[path + (bn, "scale")]))
            st = sub(stats_tree, path + (bn,))
            stats.append(torch.as_tensor(np.asarray(st["mean"], dtype=np.float32)).reshape(-1))
            stats.append(torch.as_tensor(np.asarray(st["var"], dtype=np.float32)).reshape(-1))
            self.layers.append(path + (conv,))
            return co, hout

        if W != H:
            raise ValueError("ResNet1M: square inputs expected")
        C, H = conv_bn((), "Conv_0", "BatchNorm_0", C, H, 1, cabi.OP_CONV2D, cabi.OP_BATCHNORM)
        self.ops.append(cabi.LayerDesc(cabi.OP_RELU, 0, 0, 0, 0))
        blocks = sorted((k for k in tree if k.startswith("BasicBlock_")), key=lambda k: int(k.split("_")[1]))
        for name in blocks:
            blk = tree[name]
            stride = 2 if "Conv_2" in blk else 1
            self.ops.append(cabi.LayerDesc(cabi.OP_RES_SAVE, 0, 0, 0, 0))
            c1, h1 = conv_bn((name,), "Conv_0", "BatchNorm_0", C, H, stride, cabi.OP_CONV2D, cabi.OP_BATCHNORM)
            self.ops.append(cabi.LayerDesc(cabi.OP_RELU, 0, 0, 0, 0))
            c2, h2 = conv_bn((name,), "Conv_1", "BatchNorm_1", c1, h1, 1, cabi.OP_CONV2D, cabi.OP_BATCHNORM)
            if "Conv_2" in blk:
                conv_bn((name,), "Conv_2", "BatchNorm_2", C, H, stride, cabi.OP_RES_CONV2D, cabi.OP_RES_BATCHNORM)
            self.ops.append(cabi.LayerDesc(cabi.OP_RES_ADD, 0, 0, 0, 0))
            self.ops.append(cabi.LayerDesc(cabi.OP_RELU, 0, 0, 0, 0))
            C, H = c2, h2
        dk = sub(tree, ("Dense_0", "kernel"))
        if int(dk.shape[0]) != C:
            raise ValueError(f"ResNet1M: Dense_0 kernel {tuple(dk.shape)} does not fit {C} channels")
        self.num_outputs = int(dk.shape[1])
        self.ops.append(cabi.LayerDesc(cabi.OP_GLOBAL_MEAN, 0, 0, 0, 0))
        self.ops.append(cabi.LayerDesc(cabi.OP_DENSE, C, self.num_outputs, offs[("Dense_0", "bias")], offs[("Dense_0", "kernel")]))
        self.layers.append(("Dense_0",))
        self.bn_stats = torch.cat(stats)

    def descs(self):
        arr = (cabi.LayerDesc * len(self.ops))(*self.ops)
        return arr, len(self.ops)


class BoundModel:
    """lip_model handle bound to (theta, Z): owns the activation cache; exposes the probe-batched operators."""

    def __init__(self, spec: MLPSpec, theta: torch.Tensor, Z: torch.Tensor, logvar: float = 0.0,
                 tensor_path: Optional[bool] = None):
        L = cabi.lib()
        self.spec = spec
        self.device = _require_cuda()
        self.theta = dev_f32(theta, self.device)
        if self.theta.numel() != spec.num_params:
            raise ValueError(f"flat parameter vector has {self.theta.numel()} entries, architecture needs {spec.num_params}")
        Zf = dev_f32(Z, self.device)
        self._Z_src, self._tensor_path = Zf, tensor_path
        self.M = int(Zf.shape[0])
        self.Z = Zf.reshape(self.M, -1)
        in_features = spec.in_features if hasattr(spec, "in_features") else spec.dims[0]
        if self.Z.shape[1] != in_features:
            raise ValueError(f"points have {self.Z.shape[1]} features, model expects {in_features}")
        self.D = spec.num_params
        self.K = spec.num_outputs if hasattr(spec, "num_outputs") else spec.dims[-1]
        self.model_type = spec.model_type
        self.logvar = float(logvar)
        arr, n = spec.descs()
        h = C.c_void_p()
        mt = cabi.REGRESSOR if spec.model_type == "regressor" else cabi.CLASSIFIER
        cabi.check(L.lip_model_create(arr, n, mt, self.D, C.byref(h)), "lip_model_create")
        self._h = h
        self._finalizer = weakref.finalize(self, L.lip_model_destroy, h)
        if tensor_path is not None:
            cabi.check(L.lip_model_set_tensor_path(self._h, 1 if tensor_path else 0))
        if getattr(spec, "bn_stats", None) is not None:
            self._bn_stats = dev_f32(spec.bn_stats, self.device)
            cabi.check(L.lip_model_set_bn_stats(self._h, ptr(self._bn_stats), self._bn_stats.numel(), stream()),
                       "lip_model_set_bn_stats")
        cabi.check(L.lip_model_bind(self._h, ptr(self.theta), ptr(self.Z), self.M, self.logvar, stream()),
                   "lip_model_bind")
        self._ws = _SHARED_WS
        self.launches = 0

    def clone(self) -> "BoundModel":
        """A second, independent handle bound to the same weights and points (its own activation cache, flags and side stream):
        what a second host thread / CUDA stream needs to run the operators concurrently with this one."""
        return BoundModel(self.spec, self.theta, self._Z_src, self.logvar, self._tensor_path)

    def tensor_layers(self) -> int:
        """Number of layers / conv units whose GEMMs run on the tcgen05 path (0 = everything on the fp32 SIMT kernels)."""
        return int(cabi.lib().lip_model_tensor_layers(self._h))

    def path_name(self) -> str:
        n = cabi.lib().lip_model_tensor_layers(self._h)
        total = len(self.spec.layers)
        if isinstance(self.spec, ResNetProgramSpec) and n:
            return f"implicit-GEMM conv: tcgen05-3xtf32 ({n} conv units) + simt-fp32 ({total} conv/dense stages in all)"
        if isinstance(self.spec, ConvProgramSpec):
            nf = int(cabi.lib().lip_model_fused_stages(self._h))
            if nf:
                tail = f"dense tail on tcgen05-3xtf32 ({n} layers) + simt-fp32" if n else "simt-fp32 GEMMs"
                return f"fused conv+mask+pool stage kernels ({nf} conv stages, no patch buffer) + {tail} ({total} stages in all)"
        if isinstance(self.spec, (ConvProgramSpec, ResNetProgramSpec)):
            return f"im2col + simt-fp32 ({total} conv/dense stages)"
        return f"tcgen05-3xtf32 ({n}/{total} layers) + simt-fp32" if n else "simt-fp32"

    # ---- helpers ----
    def _workspace(self, B: int):
        need = cabi.lib().lip_workspace_bytes(self._h, B)
        return self._ws.get(need)

    def outputs(self) -> torch.Tensor:
        out = torch.empty(self.M, self.K, device=self.device, dtype=torch.float32)
        cabi.check(cabi.lib().lip_model_outputs(self._h, ptr(out), stream()))
        return out

    @staticmethod
    def _check_last(t: torch.Tensor, n: int, what: str) -> None:
        if t.dim() == 0 or t.shape[-1] != n:
            raise ValueError(f"{what}: last dimension must be {n}, got shape {tuple(t.shape)}")

    # ---- operators (all probe-batched: leading dim B) ----
    def ggn_vp(self, V: torch.Tensor, recal: float, alpha: float = 0.0) -> torch.Tensor:
        exact = is_exact_tf32(V)
        V = row_strided_f32(V, self.device)
        self._check_last(V, self.D, "parameter-space vector")
        single = V.dim() == 1
        Vb = V.reshape(-1, self.D) if V.dim() != 2 else V
        B = Vb.shape[0]
        ldv = Vb.stride(0) if B > 1 else max(Vb.stride(0), self.D)
        if ldv != self.D and isinstance(self.spec, (ConvProgramSpec, ResNetProgramSpec)):
            # conv stage programs read dense [B, D] blocks (lip_ggn_vp_ex): padded probe rows are repacked once
            Vb, ldv, exact = Vb.contiguous(), self.D, False
        out = torch.empty(B, self.D, device=self.device, dtype=torch.float32)
        ws, nb = self._workspace(B)
        flags = cabi.PROBES_EXACT_TF32 if exact else 0
        cabi.check(cabi.lib().lip_ggn_vp_ex(self._h, ptr(Vb), ldv, ptr(out), self.D, B, recal, alpha, flags, ws, nb, stream()),
                   "lip_ggn_vp")
        return out.reshape(self.D) if single else out

    def wt(self, V: torch.Tensor, scale: float = 1.0, factor: int = cabi.FACTOR_SQRT) -> torch.Tensor:
        V = dev_f32(V, self.device)
        self._check_last(V, self.D, "parameter-space vector")
        single = V.dim() == 1
        Vb = V.reshape(-1, self.D)
        B = Vb.shape[0]
        out = torch.empty(B, self.M, self.K, device=self.device, dtype=torch.float32)
        ws, nb = self._workspace(B)
        cabi.check(cabi.lib().lip_wt_apply(self._h, ptr(Vb), ptr(out), B, scale, factor, ws, nb, stream()), "lip_wt_apply")
        return out[0] if single else out

    def w(self, U: torch.Tensor, scale: float = 1.0, factor: int = cabi.FACTOR_SQRT, add: Optional[torch.Tensor] = None,
          add_scale: float = 0.0, batched: Optional[bool] = None) -> torch.Tensor:
        U = dev_f32(U, self.device)
        d = self.M * self.K
        if U.numel() == 0 or U.numel() % d != 0:
            raise ValueError(f"output-space block: {tuple(U.shape)} is not a multiple of (M, K) = ({self.M}, {self.K})")
        if batched is None:
            batched = U.numel() != d
        Ub = U.reshape(-1, d)
        B = Ub.shape[0]
        out = torch.empty(B, self.D, device=self.device, dtype=torch.float32)
        addp = None
        if add is not None:
            addp = dev_f32(add, self.device).reshape(B, self.D)
        ws, nb = self._workspace(B)
        cabi.check(cabi.lib().lip_w_apply(self._h, ptr(Ub), ptr(out), B, scale, factor, ptr(addp), add_scale, ws, nb,
                                          stream()), "lip_w_apply")
        return out if batched else out.reshape(self.D)

    def gram(self, scale: float = 1.0, block: int = 256) -> torch.Tensor:
        d = self.M * self.K
        block = max(1, min(int(block), d))
        G = torch.empty(d, d, device=self.device, dtype=torch.float32)
        need = cabi.lib().lip_gram_workspace_bytes(self._h, block)
        ws, nb = self._ws.get(need)
        cabi.check(cabi.lib().lip_gram_wtw(self._h, ptr(G), scale, block, ws, nb, stream()), "lip_gram_wtw")
        return G

    def gram_cross(self, other: "BoundModel", scale_self: float = 1.0, scale_other: float = 1.0, block: int = 256) -> torch.Tensor:
        """W_self^T W_other as a [d_self, d_other] view (lip_gram_cross; build_WTWz of ggn.py:233-272)."""
        d_o = other.M * other.K
        d_s = self.M * self.K
        block = max(1, min(int(block), d_o))
        Gt = torch.empty(d_o, d_s, device=self.device, dtype=torch.float32)
        need = cabi.lib().lip_gram_cross_workspace_bytes(self._h, other._h, block)
        ws, nb = self._ws.get(need)
        cabi.check(cabi.lib().lip_gram_cross(self._h, other._h, ptr(Gt), scale_self, scale_other, block, ws, nb, stream()),
                   "lip_gram_cross")
        return Gt.T

    def zgrad(self, mode: int, X1: torch.Tensor, X2: torch.Tensor, scale: float = 1.0, per_probe: bool = False) -> torch.Tensor:
        """d/dZ of <cotangent, operator(vector)> (lip_zgrad, include/lip_b200.h): the VJP-with-respect-to-Z rule of
        ggn_vp / WTfun / Wfun / the plain batched JVP.  Returns [M, in] (summed over probes) or [B, M, in]."""
        X1 = dev_f32(X1, self.device)
        X2 = dev_f32(X2, self.device)
        self._check_last(X1, self.D, "parameter-space vector")
        X1b = X1.reshape(-1, self.D)
        B = X1b.shape[0]
        d = self.M * self.K
        if mode == cabi.ZGRAD_GGN:
            self._check_last(X2, self.D, "parameter-space vector")
            X2b = X2.reshape(-1, self.D)
        else:
            if X2.numel() != B * d:
                raise ValueError(f"output-space block: {tuple(X2.shape)} does not hold {B} x (M, K) = ({self.M}, {self.K}) entries")
            X2b = X2.reshape(B, d)
        if X2b.shape[0] != B:
            raise ValueError(f"zgrad: {B} cotangents against {X2b.shape[0]} vectors")
        in_features = self.Z.shape[1]
        out = torch.empty((B, self.M, in_features) if per_probe else (self.M, in_features), device=self.device, dtype=torch.float32)
        need = cabi.lib().lip_zgrad_workspace_bytes(self._h, mode, B)
        ws, nb = self._ws.get(need)
        cabi.check(cabi.lib().lip_zgrad(self._h, mode, ptr(X1b), ptr(X2b), ptr(out), B, scale, 1 if per_probe else 0, ws, nb,
                                        stream()), "lip_zgrad")
        return out
