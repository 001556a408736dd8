"""Manual GPU harness / test helper: LeNet5 GGN-vector products, W and W^T on fixed inputs, saved to argv[1] (.npz).  Run under
LIP_CNN_FUSE=0 / LIP_CNN_TC_TAIL=0 to exercise the im2col + GEMM stage path and the SIMT dense tail (both switches are read once per
process)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np, torch
from helpers import make_pair
from lip_b200 import ggn
ost, lst = make_pair("lenet5", seed=31)
rng = np.random.default_rng(32)
M = int(sys.argv[2]) if len(sys.argv) > 2 else 7
Z = rng.random((M, 28, 28, 1)).astype(np.float32)
D = ost.flat()[0].size
V = rng.choice([-1.0, 1.0], size=(3, D)).astype(np.float32)
V[2] = rng.standard_normal(D).astype(np.float32)
Zd, Vd = torch.as_tensor(Z, device="cuda"), torch.as_tensor(V, device="cuda")
vp = ggn.compute_ggn_vp(lst, Zd, "classifier", full_set_size=60000)
Wf, WTf = ggn.compute_W_vps(lst, Zd, "classifier", full_set_size=60000)
g = vp(Vd)
wt = WTf(Vd)
w = Wf(wt)
bm = ggn._bind(lst, Zd, "classifier")
np.savez(sys.argv[1], ggn=g.cpu().numpy(), wt=wt.cpu().numpy(), w=w.cpu().numpy())
print("path:", bm.path_name())
