"""Manual GPU harness: lip_slq_quadrature (GKL form) with the reduced / explicit u basis (LIP_GKL_REDUCED=1/0) on a small classifier."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np, torch
from helpers import make_pair
from lip_b200 import ggn, matfree, _cabi
from lip_b200._runtime import ptr, stream
kind = sys.argv[2] if len(sys.argv) > 2 else "classifier"
rng = np.random.default_rng(0)
M = int(sys.argv[1]) if len(sys.argv) > 1 else 40
if kind == "lenet5":
    ost, lst = make_pair("lenet5", seed=3)
    Z = rng.random((M, 28, 28, 1)).astype(np.float32)
elif kind == "resnet1m":
    ost, lst = make_pair("resnet1m", n_out=10, seed=3)
    Z = rng.random((M, 32, 32, 3)).astype(np.float32)
else:
    ost, lst = make_pair("classifier", hidden=(16, 16), n_out=2, in_dim=2, seed=3)
    Z = rng.standard_normal((M, 2)).astype(np.float32)
Zd = torch.as_tensor(Z, device="cuda")
Wz, WzT = ggn.compute_W_vps(lst, Zd, "classifier", full_set_size=None)
alpha = 1e-3
Av = matfree.gkl_target(WzT, Wz, alpha)
D = ost.flat()[0].size
B = 3
P = torch.as_tensor(rng.choice([-1.0, 1.0], size=(B, D)).astype(np.float32), device="cuda")
for k in [1, 2, 3, 5, 10, 30]:
    nout = D + M * Av._lip_model.K
    op = matfree._NativeOp(Av, None, B, D, nout, False, symmetric=False)
    ws, need = op.workspace(_cabi.KRYLOV_SLQ_GKL, k, B)
    out = torch.empty(B, device="cuda")
    rc = _cabi.lib().lip_slq_quadrature(op.ref(), ptr(P), D, k, B, _cabi.SLQ_GKL, matfree._FN["log"], -1.0, ptr(out), ptr(ws), need, stream())
    op.check(rc, "slq")
    print("REDUCED=%s k=%d D=%d" % (os.environ.get("LIP_GKL_REDUCED", "1"), k, D), out.cpu().numpy())
    print("JSON " + __import__("json").dumps({"k": k, "q": [float(x) for x in out.cpu().numpy()]}), flush=True)
