"""Manual GPU harness: one inducing-point training step on LeNet5 with the reference's config/scale/lenet5-2_mnist.yml sizes
(m = 200 inducing images, batch 128, st_samples 256, slq_samples 1, slq_num_matvecs 200, alpha 0.005, N = 60000)."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch
from helpers import make_pair
from lip_b200 import train_inducing, lla, utils, _cabi
ost, lst = make_pair("lenet5", seed=3)
D = ost.flat()[0].size
dev = torch.device("cuda")
rng = np.random.default_rng(4)
Z = torch.as_tensor(rng.random((200, 28, 28, 1), dtype=np.float32), device=dev)
X = torch.as_tensor(rng.random((128, 28, 28, 1), dtype=np.float32), device=dev)
L = _cabi.lib()
cvp = lla.compute_curvature_approx(lst, Z, "classifier", 0.005, full_set_size=60000)
P = torch.randint(0, 2, (256, D), device=dev).float() * 2 - 1
for rep in range(3):
    torch.cuda.synchronize(); e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); g = cvp.zgrad(P, P); e1.record(); torch.cuda.synchronize()
    print(f"lip_zgrad (LeNet5, M=200, 256 probe pairs): {e0.elapsed_time(e1):.2f} ms |dZ|={float(g.norm()):.4g}", flush=True)
opt = utils.adam(0.008)
state = opt.init(Z)
Zc = Z
for rep in range(3):
    torch.cuda.synchronize(); l0 = L.lip_launch_count(); t0 = time.perf_counter()
    Zc, state, loss = train_inducing.optimize_step(Zc, X, lst, 0.005, state, rep, opt, None, "classifier", full_set_size=60000,
                                                   scalable=True, st_samples=256, slq_samples=1, slq_num_matvecs=200)
    torch.cuda.synchronize(); dt = time.perf_counter() - t0
    print(f"optimize_step LeNet5 (m=200, |X|=128, st_samples=256, slq 1 x k=200): loss={float(loss):.6g} {dt:.3f} s "
          f"launches={L.lip_launch_count() - l0} |Z - Z0|={float((Zc - Z).norm()):.4g}", flush=True)
