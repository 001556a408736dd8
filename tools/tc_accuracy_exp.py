"""Manual GPU harness: accuracy / speed of the tcgen05 3xTF32 path vs TMEM chunk length and accumulator merging."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np, torch
from helpers import make_pair, rel_err
from oracle import lip_oracle as O
from lip_b200 import lla

ost, lst = make_pair("large", hidden=[1024, 512, 256, 128], n_out=10, in_dim=784, seed=1003, in_shape=(28, 28, 1))
rng = np.random.default_rng(1004)
Z = rng.random((512, 784)).astype(np.float32)
D = ost.flat()[0].size
V = rng.choice([-1.0, 1.0], size=(4, D)).astype(np.float32)
V[2:] = rng.standard_normal((2, D)).astype(np.float32)
ref_vp = O.compute_curvature_approx(ost, Z, "classifier", 1e-3, full_set_size=60000)
ref = np.stack([ref_vp(v) for v in V.astype(np.float64)])
cu = lambda x: torch.as_tensor(x, device="cuda")
Vb = torch.randn(128, D, device="cuda")
for tp in (False, True):
    for kc, merge in ((8, 0), (8, 1), (2, 0), (16, 0), (16, 1), (64, 0), (64, 1)):
        if not tp and (kc, merge) != (8, 0):
            continue
        os.environ["LIP_TC_KC"] = str(kc); os.environ["LIP_TC_MERGE"] = str(merge)
        cvp = lla.compute_curvature_approx(lst, cu(Z), "classifier", 1e-3, full_set_size=60000, tensor_path=tp)
        got = cvp(cu(V)).cpu().numpy()
        errs = [rel_err(got[i], ref[i]) for i in range(4)]
        maxel = [float(np.max(np.abs(got[i] - ref[i])) / np.max(np.abs(ref[i]))) for i in range(4)]
        cvp(Vb); torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(5): cvp(Vb)
        torch.cuda.synchronize()
        ms = (time.perf_counter() - t0) / 5 * 1e3
        print(f"tensor_path={tp} KC={kc} merge={merge}: rel_err {['%.2e' % e for e in errs]}  max-abs/max {['%.2e' % e for e in maxel]}  {ms:.2f} ms / 128 probes", flush=True)
