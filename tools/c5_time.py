"""Manual GPU harness: the C5 pipeline (ResNet1M, CIFAR-shaped inputs) end to end at the reference's own M = 100
(config/scale/resnet1-2_cifar10.yml:15): posterior sampler A^{-1/2} eps (Gram + LU + Lanczos 2M + W / W^T) and the sampled
predictive at 64 test points."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch
import lip_b200
from lip_b200 import lla, sample as S
from helpers import make_pair
M = int(sys.argv[1]) if len(sys.argv) > 1 else 100
Sn = int(sys.argv[2]) if len(sys.argv) > 2 else 8
ost, lst = make_pair("resnet1m", n_out=10, seed=1005, in_shape=(32, 32, 3))
rng = np.random.default_rng(1)
Z = torch.as_tensor(rng.random((M, 32, 32, 3), dtype=np.float32), device="cuda")
D = ost.flat()[0].size
Eps = torch.randn(Sn, D, device="cuda")
torch.cuda.synchronize(); t0 = time.perf_counter()
w = S.sample(lst, Z, D, 5e-3, 0, "classifier", num_samples=Sn, full_set_size=49000, eps=Eps)
torch.cuda.synchronize(); t1 = time.perf_counter()
Xnew = torch.as_tensor(rng.random((64, 32, 32, 3), dtype=np.float32), device="cuda")
pred = lla.predict_lla_scalable(lst, Xnew, Z, "classifier", 5e-3, full_set_size=49000, num_samples=Sn, eps=Eps)
torch.cuda.synchronize(); t2 = time.perf_counter()
print(f"C5 ResNet1M D={D} M={M} (d={M*10}): sample() S={Sn} {t1-t0:.2f} s (Gram d x d + LU + Lanczos-{2*M} + W/W^T); "
      f"predict_lla_scalable(64 test images, S={Sn}) incl. its own sample() {t2-t1:.2f} s; finite={bool(torch.isfinite(pred).all())}; "
      f"peak memory {torch.cuda.max_memory_allocated()/2**30:.1f} GiB", flush=True)
