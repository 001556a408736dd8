"""LeNet5's three execution paths agree: fused conv stages + tcgen05 dense tail (default), fused stages + SIMT tail
(LIP_CNN_TC_TAIL=0), im2col + GEMM stages (LIP_CNN_FUSE=0; the path other conv geometries take).  The switches are read once per
process, so each path runs in its own interpreter (tools/lenet_paths_check.py); the default path is pinned on the oracle in
test_gpu_parity.py::test_lenet5_operators_match_oracle."""
import os
import subprocess
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(tmp_path, name, **env):
    out = str(tmp_path / f"{name}.npz")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "lenet_paths_check.py"), out, "18"], env=dict(os.environ, **env),
                       capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    return np.load(out), r.stdout


def test_lenet5_paths_agree(tmp_path):
    ref, log0 = _run(tmp_path, "default")
    assert "fused conv" in log0 and "tcgen05" in log0, log0
    simt_tail, log1 = _run(tmp_path, "simt_tail", LIP_CNN_TC_TAIL="0")
    assert "fused conv" in log1 and "tcgen05" not in log1, log1
    unfused, log2 = _run(tmp_path, "unfused", LIP_CNN_FUSE="0", LIP_CNN_TC_TAIL="0")
    assert "im2col" in log2, log2
    for other in (simt_tail, unfused):
        for key in ("ggn", "wt", "w"):
            a, b = other[key].astype(np.float64), ref[key].astype(np.float64)
            assert np.linalg.norm(a - b) <= 2e-5 * np.linalg.norm(b), key
