"""The GKL logdet quadrature with the u basis in reduced coordinates (csrc/lip_krylov.cu gkl_run, the default of lip_slq_quadrature
for the structured operator) against the explicit recurrence (LIP_GKL_REDUCED=0) on the conv models' operators.  The switch is read
once per process, so each form runs in its own interpreter (tools/debug_gkl_reduced.py)."""
import json
import os
import subprocess
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(kind, reduced):
    env = dict(os.environ, LIP_GKL_REDUCED="1" if reduced else "0")
    out = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "debug_gkl_reduced.py"), "6", kind], env=env, capture_output=True,
                         text=True, timeout=600)
    assert out.returncode == 0, out.stderr[-2000:]
    return {r["k"]: np.array(r["q"]) for r in (json.loads(l[5:]) for l in out.stdout.splitlines() if l.startswith("JSON "))}


@pytest.mark.parametrize("kind", ["lenet5", "classifier"])
def test_reduced_u_basis_equals_explicit_recurrence(kind):
    a, b = _run(kind, True), _run(kind, False)
    assert sorted(a) == sorted(b) and len(a) >= 5
    for k in a:
        rel = np.abs(a[k] - b[k]) / np.abs(b[k])
        assert np.all(rel < 2e-5), (kind, k, rel)
