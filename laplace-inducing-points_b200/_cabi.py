"""ctypes binding of include/lip_b200.h (the drop-in C ABI).  There is NO fallback: if the CUDA library is
missing or a call fails, this raises."""
from __future__ import annotations

import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "liblip_b200.so")

LIP_OK = 0
ERR_INVALID, ERR_CUDA, ERR_WORKSPACE, ERR_UNSUPPORTED, ERR_NOT_BOUND = -1, -2, -3, -4, -5
OP_DENSE, OP_TANH, OP_GELU_TANH, OP_RELU = 0, 1, 2, 3
OP_CONV2D, OP_AVGPOOL2, OP_ZEROPAD, OP_FLATTEN, OP_INPUT = 4, 5, 6, 7, 8
OP_BATCHNORM, OP_RES_SAVE, OP_RES_CONV2D, OP_RES_BATCHNORM, OP_RES_ADD, OP_GLOBAL_MEAN = 9, 10, 11, 12, 13, 14
REGRESSOR, CLASSIFIER = 0, 1
FACTOR_NONE, FACTOR_SQRT = 0, 1
ZGRAD_GGN, ZGRAD_WT, ZGRAD_W, ZGRAD_JVP = 0, 1, 2, 3
FN_LOG, FN_INVSQRT, FN_INV, FN_IDENTITY, FN_SAMPLER = 0, 1, 2, 3, 4
LINOP_GGN, LINOP_GKL, LINOP_DENSE_SYM, LINOP_CALLBACK = 0, 1, 2, 3
KRYLOV_LANCZOS, KRYLOV_GKL, KRYLOV_SLQ_LANCZOS, KRYLOV_SLQ_GKL, KRYLOV_FUNM, KRYLOV_CG, KRYLOV_HUTCHPP, KRYLOV_APPLY = 0, 1, 2, 3, 4, 5, 6, 7
SLQ_LANCZOS, SLQ_GKL = 0, 1
PROBES_EXACT_TF32 = 1


class LayerDesc(C.Structure):
    _fields_ = [("op", C.c_int32), ("in_features", C.c_int32), ("out_features", C.c_int32),
                ("bias_offset", C.c_int64), ("kernel_offset", C.c_int64),
                ("kh", C.c_int32), ("kw", C.c_int32), ("stride", C.c_int32), ("pad", C.c_int32)]


_P, _I32, _I64, _F, _SZ = C.c_void_p, C.c_int32, C.c_int64, C.c_float, C.c_size_t

# lip_matvec_fn: int fn(void* ctx, int32_t transpose, int64_t B, lip_stream_t stream)
MATVEC_FN = C.CFUNCTYPE(C.c_int, C.c_void_p, C.c_int32, C.c_int64, C.c_void_p)


class LinOp(C.Structure):
    """lip_linop of include/lip_b200.h"""
    _fields_ = [("kind", C.c_int32), ("symmetric", C.c_int32), ("model", C.c_void_p),
                ("scale", C.c_float), ("alpha", C.c_float), ("beta", C.c_float),
                ("dense", C.c_void_p), ("n", C.c_int64), ("n_out", C.c_int64),
                ("fn", MATVEC_FN), ("ctx", C.c_void_p),
                ("cb_in", C.c_void_p), ("cb_out", C.c_void_p), ("cb_in_t", C.c_void_p), ("cb_out_t", C.c_void_p)]


_LP = C.POINTER(LinOp)

# name -> (restype, argtypes); every symbol declared in include/lip_b200.h
SIGNATURES = {
    "lip_last_error": (C.c_char_p, []),
    "lip_version": (C.c_int, []),
    "lip_launch_count": (_I64, []),
    "lip_device_is_sm100": (C.c_int, []),
    "lip_model_create": (C.c_int, [C.POINTER(LayerDesc), _I32, _I32, _I64, C.POINTER(_P)]),
    "lip_model_destroy": (C.c_int, [_P]),
    "lip_model_num_params": (_I64, [_P]),
    "lip_model_num_outputs": (_I64, [_P]),
    "lip_model_num_points": (_I64, [_P]),
    "lip_model_set_tensor_path": (C.c_int, [_P, _I32]),
    "lip_model_tensor_layers": (C.c_int, [_P]),
    "lip_model_fused_stages": (C.c_int, [_P]),
    "lip_model_bind": (C.c_int, [_P, _P, _P, _I64, _F, _P]),
    "lip_model_set_bn_stats": (C.c_int, [_P, _P, _I64, _P]),
    "lip_model_outputs": (C.c_int, [_P, _P, _P]),
    "lip_workspace_bytes": (_SZ, [_P, _I64]),
    "lip_ggn_vp": (C.c_int, [_P, _P, _P, _I64, _F, _F, _P, _SZ, _P]),
    "lip_ggn_vp_ex": (C.c_int, [_P, _P, _I64, _P, _I64, _I64, _F, _F, _I32, _P, _SZ, _P]),
    "lip_wt_apply": (C.c_int, [_P, _P, _P, _I64, _F, _I32, _P, _SZ, _P]),
    "lip_w_apply": (C.c_int, [_P, _P, _P, _I64, _F, _I32, _P, _F, _P, _SZ, _P]),
    "lip_gram_wtw": (C.c_int, [_P, _P, _F, _I64, _P, _SZ, _P]),
    "lip_gram_workspace_bytes": (_SZ, [_P, _I64]),
    "lip_gram_cross": (C.c_int, [_P, _P, _P, _F, _F, _I64, _P, _SZ, _P]),
    "lip_gram_cross_workspace_bytes": (_SZ, [_P, _P, _I64]),
    "lip_zgrad": (C.c_int, [_P, _I32, _P, _P, _P, _I64, _F, _I32, _P, _SZ, _P]),
    "lip_zgrad_workspace_bytes": (_SZ, [_P, _I32, _I64]),
    "lip_mc_softmax_predictive": (C.c_int, [_P, _P, _P, _P, _I64, _I64, _I32, _P]),
    "lip_dot_scratch_bytes": (_SZ, [_I64, _I64]),
    "lip_dot": (C.c_int, [_P, _P, _P, _I64, _I64, _I64, _I64, _P, _P]),
    "lip_axpby": (C.c_int, [_P, _P, _P, _P, _I64, _I64, _I64, _I64, _P]),
    "lip_scale": (C.c_int, [_P, _I32, _P, _P, _I64, _I64, _I64, _I64, _P]),
    "lip_unpack_rademacher": (C.c_int, [_P, _I64, _P, _I64, _I64, _P]),
    "lip_unpack_rademacher_ld": (C.c_int, [_P, _I64, _P, _I64, _I64, _I64, _P]),
    "lip_cg_step": (C.c_int, [_P, _P, _P, _P, _P, _P, _P, _P, _I64, _I64, _P, _P]),
    "lip_cg_init": (C.c_int, [_P, _P, _P, _P, _P, _P, _P, _P, _F, _F, _I64, _I64, _P, _P]),
    "lip_reorth_scratch_bytes": (_SZ, [_I64, _I64, _I64]),
    "lip_reorth": (C.c_int, [_P, _I64, _I64, _I64, _P, _I64, _P, _P, _I32, _I64, _I64, _P, _P]),
    "lip_basis_combine": (C.c_int, [_P, _I64, _I64, _I64, _P, _I64, _P, _I64, _I64, _I64, _P]),
    "lip_tridiag_scratch_bytes": (_SZ, [_I64, _I64, _I32]),
    "lip_tridiag_funm": (C.c_int, [_P, _P, _I64, _I64, _I32, _F, _P, _P, _P, _P, _P]),
    "lip_tridiag_funm_p": (C.c_int, [_P, _P, _I64, _I64, _I32, _F, C.POINTER(C.c_float), _P, _P, _P, _P, _P]),
    "lip_bidiag_to_tridiag": (C.c_int, [_P, _P, _P, _P, _I64, _I64, _P]),
    "lip_krylov_workspace_bytes": (_SZ, [_LP, _I32, _I64, _I64]),
    "lip_linop_apply": (C.c_int, [_LP, _P, _P, _I64, _I32, _P, _SZ, _P]),
    "lip_lanczos_tridiag": (C.c_int, [_LP, _P, _I64, _I64, _I64, _I32, _P, _I64, _P, _P, _P, _P, _SZ, _P]),
    "lip_gkl_bidiag": (C.c_int, [_LP, _P, _I64, _I64, _I64, _P, _I64, _P, _I64, _P, _P, _P, _P, _SZ, _P]),
    "lip_slq_quadrature": (C.c_int, [_LP, _P, _I64, _I64, _I64, _I32, _I32, _F, _P, _P, _SZ, _P]),
    "lip_comm_unique_id": (C.c_int, [_P]),
    "lip_comm_create": (C.c_int, [_P, _I32, _I32, C.POINTER(_P)]),
    "lip_comm_destroy": (C.c_int, [_P]),
    "lip_comm_world": (C.c_int, [_P]),
    "lip_comm_rank": (C.c_int, [_P]),
    "lip_comm_allreduce_sum": (C.c_int, [_P, _P, _I64, _P]),
    "lip_slq_quadrature_sharded": (C.c_int, [_LP, _P, _P, _I64, _I64, _I64, _I32, _I32, _F, _P, _P, _SZ, _P]),
    "lip_slq_workspace_bytes": (_SZ, [_LP, _I32, _I64, _I64, _I32]),
    "lip_funm_lanczos": (C.c_int, [_LP, _P, _I64, _I64, _I64, _I32, _F, C.POINTER(C.c_float), _P, _I64, _P, _SZ, _P]),
    "lip_hutchpp_v2": (C.c_int, [_LP, _P, _I64, _I64, _I64, _P, _P, _P, _SZ, _P]),
    "lip_cg_solve": (C.c_int, [_LP, _P, _P, _I64, _F, _F, _I64, _I32, _P, _P, _SZ, _P]),
    "lip_bench_tc_gemm": (C.c_int, [_I32, _I64, _I64, _I64, _I64, _I32, _I32, _I32, C.POINTER(_F), _P]),
    "lip_selftest_tc_gemm": (C.c_int, [_I32, _I64, _I64, _I64, _I64, C.POINTER(_F), _P]),
    "lip_selftest_conv_tc": (C.c_int, [_I32, _I64, _I32, _I32, _I32, _I32, _I32, _I32, _I64, _I32, C.POINTER(_F), C.POINTER(_F),
                                       C.POINTER(_F), _P]),
}

_lib = None


class LipError(RuntimeError):
    pass


def lib() -> C.CDLL:
    """Loads liblip_b200.so; raises if it has not been built (python __graft_entry__.py build)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise LipError(f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                           "(nvcc, sm_100a).  There is no CPU or PyTorch fallback for this path.")
        L = C.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(L, name)  # AttributeError if the symbol is not exported
            fn.restype = res
            fn.argtypes = args
        _lib = L
    return _lib


def check(rc: int, what: str = "") -> None:
    if rc == LIP_OK:
        return
    msg = lib().lip_last_error().decode("utf-8", "replace")
    text = f"{what}: {msg}" if what else msg
    if rc in (ERR_INVALID, ERR_WORKSPACE, ERR_NOT_BOUND):
        raise ValueError(text)
    raise LipError(text)
