"""Manual GPU harness: one inducing-point training step on ResNet1M with the reference's config/scale/resnet1-2_cifar10.yml sizes
(m = 100 inducing images, batch 32, st_samples 24, slq_samples 1, slq_num_matvecs 16, alpha 0.005, N = 49000), and lip_zgrad alone."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch
from helpers import make_pair
from lip_b200 import train_inducing, lla, utils, _cabi
ost, lst = make_pair("resnet1m", n_out=10, in_shape=(32, 32, 3), seed=3)
D = ost.flat()[0].size
dev = torch.device("cuda")
rng = np.random.default_rng(4)
Z = torch.as_tensor(rng.random((100, 32, 32, 3), dtype=np.float32), device=dev)
X = torch.as_tensor(rng.random((32, 32, 32, 3), dtype=np.float32), device=dev)
L = _cabi.lib()
cvp = lla.compute_curvature_approx(lst, Z, "classifier", 0.005, full_set_size=49000)
P = torch.randint(0, 2, (24, D), device=dev).float() * 2 - 1
for rep in range(3):
    torch.cuda.synchronize(); e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    l0 = L.lip_launch_count()
    e0.record(); g = cvp.zgrad(P, P); e1.record(); torch.cuda.synchronize()
    print(f"lip_zgrad (ResNet1M, M=100, 24 probe pairs): {e0.elapsed_time(e1):.2f} ms launches={L.lip_launch_count() - l0} |dZ|={float(g.norm()):.4g} finite={bool(torch.isfinite(g).all())}", flush=True)
opt = utils.adam(0.005)
state = opt.init(Z)
Zc = Z
for rep in range(3):
    torch.cuda.synchronize(); l0 = L.lip_launch_count(); t0 = time.perf_counter()
    Zc, state, loss = train_inducing.optimize_step(Zc, X, lst, 0.005, state, rep, opt, None, "classifier", full_set_size=49000,
                                                   scalable=True, st_samples=24, slq_samples=1, slq_num_matvecs=16)
    torch.cuda.synchronize(); dt = time.perf_counter() - t0
    print(f"optimize_step ResNet1M (m=100, |X|=32, st_samples=24, slq 1 x k=16): loss={float(loss):.6g} {dt:.3f} s "
          f"launches={L.lip_launch_count() - l0} |Z - Z0|={float((Zc - Z).norm()):.4g}", flush=True)
