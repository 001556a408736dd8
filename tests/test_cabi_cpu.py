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
