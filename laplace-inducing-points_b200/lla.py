"""Drop-in mirror of the hot-path part of /root/reference/src/lla.py on the B200 path."""
from __future__ import annotations

import torch

from . import _cabi as cabi
from ._runtime import dev_f32
from .ggn import _batched, _bind, compute_ggn_vp
from .sample import sample
from .utils import flatten_nn_params


def compute_curvature_approx(map_state, Z, model_type, alpha, full_set_size=None, *, tensor_path=None,
                             shard_points=False):
    """lla.py:11-23: curvature_vp(v) = ggn_vp(v) + alpha v — one fused call (the +alpha v is the GEMM epilogue).
    shard_points=True: points sharded over ranks (see ggn.compute_ggn_vp); each rank adds alpha / world_size of v so
    the all-reduced sum carries alpha v once."""
    ggn_vp = compute_ggn_vp(map_state, Z, model_type=model_type, full_set_size=full_set_size, tensor_path=tensor_path,
                            shard_points=shard_points)
    bm, recal = ggn_vp._lip_model, ggn_vp._lip_recal
    alpha = float(alpha)
    if shard_points:
        from . import _dist
        local_alpha = alpha / _dist.world()[1]
        fn = _dist.point_sharded(_batched(lambda v: bm.ggn_vp(v, recal, local_alpha), bm, _lip_kind="GGN"))
        fn._lip_recal, fn._lip_alpha, fn._lip_transpose = recal, alpha, fn
        return fn

    def curvature_vp(v):
        return bm.ggn_vp(v, recal, alpha)

    # d/dZ <ubar, curvature_vp(v)>: the alpha v term does not depend on Z
    return _batched(curvature_vp, bm, _lip_kind="GGN", _lip_recal=recal, _lip_alpha=alpha, _lip_transpose=curvature_vp,
                    zgrad=ggn_vp.zgrad)


def predict_lla_scalable(map_state, Xnew, Z, model_type, alpha, key=None, full_set_size=None, num_samples=1, *,
                         eps=None, sampler=None):
    """lla.py:133-156: f(theta*, X) + J_X w_s for posterior samples w_s = A^{-1/2} eps_s; returns [S, Bt, K].
    The S sequential batch-JVPs of lla.py:153-154 are one lip_wt_apply(factor=NONE) over all samples.
    `sampler` (an inv_matsqrt_vp closure for the same state / Z / alpha) lets a caller that predicts on many test batches
    (evaluate.eval_dataset) build the Gram factorisation once instead of once per batch (sample.py:77 inside evaluate.py:98-111)."""
    flat_params, _ = flatten_nn_params(map_state.params)
    D = flat_params.numel()
    key = key if key is not None else 123
    if sampler is not None:
        if eps is None:
            from .matfree import _generator
            g = _generator(key)
            eps = torch.randn(num_samples, D, generator=g, device=g.device)
        w_samples = sampler(dev_f32(eps).reshape(-1, D))
    else:
        w_samples = sample(map_state, Z, D, alpha=alpha, key=key, model_type=model_type, num_samples=num_samples,
                           full_set_size=full_set_size, eps=eps)
    bx = _bind(map_state, Xnew, model_type)
    fmu = bx.outputs()
    dys = bx.wt(w_samples.reshape(-1, D), scale=1.0, factor=cabi.FACTOR_NONE)
    return fmu[None] + dys


def materialize_covariance(f_cov_vp, N, out_dim, mode="diag"):
    """lla.py:160-217: diagonal / full matrix of an operator by unit-vector probing (float64 buffers)."""
    K = N * out_dim
    dev = torch.device("cuda", torch.cuda.current_device())
    eye = torch.eye(K, device=dev, dtype=torch.float64)
    if mode == "diag":
        diag = torch.zeros(K, device=dev, dtype=torch.float64)
        for i in range(K):
            diag[i] = torch.as_tensor(f_cov_vp(eye[i])).reshape(K)[i]
        return diag.reshape(N, out_dim)
    if mode == "full":
        cov = torch.zeros(K, K, device=dev, dtype=torch.float64)
        for i in range(K):
            cov[:, i] = torch.as_tensor(f_cov_vp(eye[i])).reshape(K)
        return cov
    raise ValueError("mode must be 'diag' or 'full'")
