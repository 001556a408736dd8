"""Multi-GPU check (run under torchrun, one rank per GPU): point-sharded and probe-sharded operators against the
single-GPU result on the same inputs.  Prints one line per check from rank 0; exits non-zero on a mismatch."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch
import __graft_entry__ as entry
entry.build()
import lip_b200
from lip_b200 import _dist, lla, ggn
from helpers import make_pair, rel_err

rank, world, local = _dist.init_from_env()
torch.cuda.set_device(local)
ok = True
for name, kw, zshape in (("mlp", dict(kind="large", hidden=[256, 128], n_out=10, in_dim=64, seed=5), (96, 64)),
                         ("lenet5", dict(kind="lenet5", seed=6), (6, 28, 28, 1))):
    kind = kw.pop("kind")
    ost, lst = make_pair(kind, **kw)
    rng = np.random.default_rng(11)                               # identical inputs on every rank
    Z = torch.as_tensor(rng.random(zshape).astype(np.float32), device="cuda")
    D = ost.flat()[0].size
    V = torch.as_tensor(rng.standard_normal((8, D)).astype(np.float32), device="cuda")
    full = lla.compute_curvature_approx(lst, Z, "classifier", 0.3, full_set_size=5000)(V)
    shard = lla.compute_curvature_approx(lst, Z, "classifier", 0.3, full_set_size=5000, shard_points=True)(V)
    e1 = rel_err(shard.cpu().numpy(), full.cpu().numpy())
    cvp = lla.compute_curvature_approx(lst, Z, "classifier", 0.3, full_set_size=5000)
    tr_full = float(((V * cvp(V)).sum(1)).mean())
    tr_shard = float(_dist.hutchinson_sharded(cvp, V))
    e2 = abs(tr_shard - tr_full) / abs(tr_full)
    if rank == 0:
        print(f"{name}: world={world} point-sharded curvature_vp rel_err={e1:.2e}  probe-sharded Hutchinson rel_err={e2:.2e}", flush=True)
    ok = ok and e1 < 1e-5 and e2 < 1e-5
    if name == "mlp":   # probe-sharded Z-gradient (SURVEY 8 rows e / f1): one all-reduce of the [M, in] result
        U = torch.as_tensor(rng.standard_normal((8, D)).astype(np.float32), device="cuda")
        g_full = cvp.zgrad(U, V)
        g_shard = _dist.zgrad_sharded(cvp.zgrad, U, V)
        e3 = rel_err(g_shard.cpu().numpy(), g_full.cpu().numpy())
        if rank == 0:
            print(f"{name}: world={world} probe-sharded lip_zgrad rel_err={e3:.2e}", flush=True)
        ok = ok and e3 < 1e-5
if world > 1:
    torch.distributed.barrier()
    torch.distributed.destroy_process_group()
sys.exit(0 if ok else 1)
