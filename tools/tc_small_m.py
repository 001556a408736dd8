"""Manual GPU harness: does the tcgen05 path hold parity / beat the SIMT kernels when fewer than 64 points are bound
(the reference's mlp_mnist.yml trains m = 50 inducing points)?  Run with LIP_TC_MIN_M=<n> in the environment."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch
import bench
from lip_b200 import ggn
for M in (16, 24, 50, 63):
    bench.M_POINTS = M
    ost, lst, Z = bench.build_states()
    D = ost.flat()[0].size
    dev = torch.device("cuda")
    Zt = torch.as_tensor(Z, device=dev)
    rng = np.random.default_rng(M)
    V = torch.as_tensor(rng.standard_normal((64, D)).astype(np.float32), device=dev)
    U = torch.as_tensor(rng.standard_normal((64, M, 10)).astype(np.float32), device=dev)
    res = {}
    for tp in (False, True):
        vp = ggn.compute_ggn_vp(lst, Zt, "classifier", full_set_size=60000, tensor_path=tp)
        Wf, WTf = ggn.compute_W_vps(lst, Zt, "classifier", tensor_path=tp)
        outs = (vp(V), Wf(U), WTf(V))
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(5): vp(V)
        e1.record(); torch.cuda.synchronize()
        res[tp] = (outs, e0.elapsed_time(e1) / 5, vp._lip_model.tensor_layers())
    errs = [float((a.double() - b.double()).norm() / b.double().norm()) for a, b in zip(res[True][0], res[False][0])]
    print(f"M={M}: tensor layers {res[True][2]}  ggn_vp 64 probes: simt {res[False][1]:.2f} ms, tc {res[True][1]:.2f} ms; "
          f"rel diff tc vs simt: ggn {errs[0]:.2e} W {errs[1]:.2e} WT {errs[2]:.2e}", flush=True)
