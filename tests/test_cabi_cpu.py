"""CPU-side checks of the drop-in boundary: the C-ABI library loads, exports every symbol include/lip_b200.h
declares, and validates its arguments (no compute calls: there is no GPU here)."""
import ctypes as C
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def cabi():
    import __graft_entry__ as g
    g.build()
    import lip_b200  # noqa: F401
    from lip_b200 import _cabi
    return _cabi


def _declared_symbols():
    src = open(os.path.join(ROOT, "include", "lip_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(lip_[a-z0-9_]+)\s*\(", src)))


def test_header_symbols_all_exported_and_bound(cabi):
    names = _declared_symbols()
    assert len(names) >= 25
    lib = cabi.lib()
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/lip_b200.h but not exported"
        assert n in cabi.SIGNATURES, f"{n} has no ctypes signature"
    assert sorted(cabi.SIGNATURES) == names


def test_model_create_validates_programs(cabi):
    lib = cabi.lib()
    h = C.c_void_p()
    good = (cabi.LayerDesc * 3)(cabi.LayerDesc(cabi.OP_DENSE, 2, 4, 0, 4), cabi.LayerDesc(cabi.OP_TANH, 0, 0, 0, 0),
                                cabi.LayerDesc(cabi.OP_DENSE, 4, 3, 12, 15))
    assert lib.lip_model_create(good, 3, cabi.CLASSIFIER, 27, C.byref(h)) == 0
    assert lib.lip_model_num_params(h) == 27 and lib.lip_model_num_outputs(h) == 3
    assert lib.lip_model_num_points(h) == -1          # not bound
    assert lib.lip_workspace_bytes(h, 8) == 0
    assert lib.lip_model_destroy(h) == 0
    # wrong parameter count
    assert lib.lip_model_create(good, 3, cabi.CLASSIFIER, 28, C.byref(h)) == cabi.ERR_INVALID
    assert b"27" in lib.lip_last_error()
    # mismatched widths
    bad = (cabi.LayerDesc * 2)(cabi.LayerDesc(cabi.OP_DENSE, 2, 4, 0, 4), cabi.LayerDesc(cabi.OP_DENSE, 5, 3, 12, 15))
    assert lib.lip_model_create(bad, 2, cabi.CLASSIFIER, 30, C.byref(h)) == cabi.ERR_INVALID
    # activation first / unknown op / regressor with K != 1
    bad2 = (cabi.LayerDesc * 1)(cabi.LayerDesc(cabi.OP_TANH, 0, 0, 0, 0))
    assert lib.lip_model_create(bad2, 1, cabi.CLASSIFIER, 0, C.byref(h)) == cabi.ERR_INVALID
    bad3 = (cabi.LayerDesc * 1)(cabi.LayerDesc(17, 2, 2, 0, 2))
    assert lib.lip_model_create(bad3, 1, cabi.CLASSIFIER, 6, C.byref(h)) == cabi.ERR_INVALID
    reg = (cabi.LayerDesc * 1)(cabi.LayerDesc(cabi.OP_DENSE, 2, 2, 0, 2))
    assert lib.lip_model_create(reg, 1, cabi.REGRESSOR, 6, C.byref(h)) == cabi.ERR_INVALID
    with pytest.raises(ValueError):
        cabi.check(cabi.ERR_INVALID, "x")


def test_argument_checks_do_not_need_a_gpu(cabi):
    lib = cabi.lib()
    assert lib.lip_dot(None, None, None, 4, 1, 4, 4, None, None) == cabi.ERR_INVALID
    assert lib.lip_tridiag_funm(None, None, 4, 1, 0, -1.0, None, None, None, None, None) == cabi.ERR_INVALID
    assert lib.lip_reorth(None, 4, 0, 1, None, 4, None, None, 2, 4, 1, None, None) == cabi.ERR_INVALID
    assert lib.lip_version() >= 100
    assert lib.lip_tridiag_scratch_bytes(10, 2, 1) >= 8 * 2 * (40 + 100)


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "laplace-inducing-points_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                text = open(os.path.join(dirpath, f)).read()
                assert "import oracle" not in text and "from oracle" not in text, f


def test_host_spec_and_layout_without_gpu():
    import lip_b200  # noqa: F401
    from lip_b200._runtime import MLPSpec
    from lip_b200 import _cabi as cb
    spec = MLPSpec([784, 1024, 512, 256, 128, 10], cb.OP_TANH, "classifier")
    assert spec.num_params == 1494154
    nin, nout, boff, woff = spec.layers[1]
    assert (nin, nout, boff, woff) == (1024, 512, 1024 + 784 * 1024, 1024 + 784 * 1024 + 512)
    arr, n = spec.descs()
    assert n == 9 and arr[1].op == cb.OP_TANH and arr[8].op == cb.OP_DENSE


def _lenet_program(cabi):
    import lip_b200  # noqa: F401
    from lip_b200._runtime import ConvProgramSpec
    ops = [("input", 28, 28, 1), ("pad", 2),
           ("conv", "Conv_0", 5, 5, 1, 6), ("act", cabi.OP_RELU), ("pool",),
           ("conv", "Conv_1", 5, 5, 6, 16), ("act", cabi.OP_RELU), ("pool",), ("flatten",),
           ("dense", "Dense_0", 400, 120), ("act", cabi.OP_RELU), ("dense", "Dense_1", 120, 84), ("act", cabi.OP_RELU),
           ("dense", "Dense_2", 84, 10)]
    return ConvProgramSpec(ops, "classifier", name="LeNet5")


def test_conv_program_create_and_layout(cabi):
    """LeNet5 (scalemodels.py:11-49) as a conv stage program: accepted by lip_model_create (host-only call), D = 61706,
    and the module offsets follow the reference's flat order (utils.py:12-17), checked against the oracle's flatten."""
    import ctypes as C
    from oracle import models as OM
    L = cabi.lib()
    spec = _lenet_program(cabi)
    assert spec.num_params == 61706 and spec.in_features == 784 and spec.num_outputs == 10
    arr, n = spec.descs()
    h = C.c_void_p()
    assert L.lip_model_create(arr, n, cabi.CLASSIFIER, spec.num_params, C.byref(h)) == 0, L.lip_last_error()
    assert L.lip_model_num_params(h) == 61706 and L.lip_model_num_outputs(h) == 10
    assert L.lip_model_destroy(h) == 0
    # offsets vs the oracle's ravel order: fill each leaf with a distinct constant and look it up in the flat vector
    om = OM.LeNet5()
    variables = om.init(0)
    code = {}
    for i, mod in enumerate(sorted(variables["params"])):
        for j, leaf in enumerate(sorted(variables["params"][mod])):
            variables["params"][mod][leaf][...] = 10 * i + j
            code[(mod, leaf)] = 10 * i + j
    flat, _ = OM.flatten_nn_params(variables["params"])
    for mod, (boff, woff) in spec.offsets.items():
        assert flat[boff] == code[(mod, "bias")] and flat[woff] == code[(mod, "kernel")]
        assert flat[woff - 1] == code[(mod, "bias")]          # bias block ends where the kernel block starts
    # malformed programs are rejected with LIP_ERR_INVALID
    def create(descs, D):
        a = (cabi.LayerDesc * len(descs))(*descs)
        hh = C.c_void_p()
        return L.lip_model_create(a, len(descs), cabi.CLASSIFIER, D, C.byref(hh))
    inp = cabi.LayerDesc(cabi.OP_INPUT, 1, 0, 0, 0, 8, 8, 0, 0)
    conv = cabi.LayerDesc(cabi.OP_CONV2D, 1, 2, 0, 2, 3, 3, 1, 0)          # 8x8x1 -> 6x6x2, 18 + 2 params
    relu = cabi.LayerDesc(cabi.OP_RELU, 0, 0, 0, 0)
    pool = cabi.LayerDesc(cabi.OP_AVGPOOL2, 0, 0, 0, 0)
    dense = cabi.LayerDesc(cabi.OP_DENSE, 18, 3, 20, 23)                   # 3x3x2 = 18 -> 3
    assert create([inp, conv, relu, pool, dense], 20 + 3 + 54) == 0
    assert create([inp, conv, relu, pool, dense], 99) == cabi.ERR_INVALID                  # parameter count mismatch
    assert create([inp, conv, pool, dense], 77) == cabi.ERR_INVALID                        # conv stage without activation
    bad_stride = cabi.LayerDesc(cabi.OP_CONV2D, 1, 2, 0, 2, 3, 3, 2, 0)
    assert create([inp, bad_stride, relu, pool, dense], 77) == cabi.ERR_INVALID            # stride 2 not built
    bad_cin = cabi.LayerDesc(cabi.OP_CONV2D, 3, 2, 0, 2, 3, 3, 1, 0)
    assert create([inp, bad_cin, relu, pool, dense], 77) == cabi.ERR_INVALID
    assert create([inp, conv, relu, pool], 20) == cabi.ERR_INVALID                         # must end with DENSE
    odd = cabi.LayerDesc(cabi.OP_INPUT, 1, 0, 0, 0, 7, 7, 0, 0)                            # 7x7 -> 5x5: odd pool
    assert create([odd, conv, relu, pool, dense], 77) == cabi.ERR_INVALID


def test_resnet_program_create(cabi):
    """ResNet1M (scalemodels.py:70-157) as a residual conv program built from the parameter tree: accepted by
    lip_model_create, D = 1,084,586, 21 conv+BatchNorm units -> 2 * 2016 running statistics."""
    import ctypes as C
    from oracle import models as OM
    from lip_b200._runtime import ResNetProgramSpec
    L = cabi.lib()
    v = OM.ResNet1M(10, (32, 32, 3)).init(0)
    spec = ResNetProgramSpec(v["params"], v["batch_stats"], (32, 32, 3), "classifier")
    assert spec.num_params == 1084586 and spec.num_outputs == 10
    assert spec.bn_stats.numel() == 2 * (32 * 7 + 64 * 7 + 128 * 7)
    arr, n = spec.descs()
    h = C.c_void_p()
    assert L.lip_model_create(arr, n, cabi.CLASSIFIER, spec.num_params, C.byref(h)) == 0, L.lip_last_error()
    assert L.lip_model_num_params(h) == 1084586 and L.lip_model_num_outputs(h) == 10
    assert L.lip_model_destroy(h) == 0
    # a block that changes shape without a shortcut conv is rejected
    bad = [d for d in spec.ops if d.op not in (cabi.OP_RES_CONV2D, cabi.OP_RES_BATCHNORM)]
    a = (cabi.LayerDesc * len(bad))(*bad)
    assert L.lip_model_create(a, len(bad), cabi.CLASSIFIER, spec.num_params, C.byref(h)) == cabi.ERR_INVALID
    # missing BatchNorm statistics are a host-side error
    import pytest
    with pytest.raises(ValueError):
        ResNetProgramSpec(v["params"], {}, (32, 32, 3), "classifier")
