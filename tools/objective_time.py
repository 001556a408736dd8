"""Manual GPU harness: one forward evaluation of the scalable KL objective (train_inducing.py:87-173) on the headline
MNIST-MLP shape, with the reference config's sizes (config/scale/mlp_mnist.yml: m, batch_size, st_samples)."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch
import bench
from lip_b200 import train_inducing, _cabi
M = int(sys.argv[1]) if len(sys.argv) > 1 else 512
batch = int(sys.argv[2]) if len(sys.argv) > 2 else 512
st_samples = int(sys.argv[3]) if len(sys.argv) > 3 else 64
k = int(sys.argv[4]) if len(sys.argv) > 4 else int(0.8 * M)
bench.M_POINTS = M
ost, lst, Z = bench.build_states()
rng = np.random.default_rng(9)
X = rng.random((batch, 784), dtype=np.float32)
D = ost.flat()[0].size
dev = torch.device("cuda")
probes = torch.randint(0, 2, (st_samples, D), device=dev).float() * 2 - 1
L = _cabi.lib()
for rep in range(2):
    torch.cuda.synchronize(); l0 = L.lip_launch_count(); t0 = time.perf_counter()
    val = float(train_inducing.alternative_objective_scalable(torch.as_tensor(Z, device=dev), torch.as_tensor(X, device=dev), lst,
                                                              bench.ALPHA, "classifier", 0, full_set_size=bench.N_FULL,
                                                              st_samples=st_samples, slq_samples=2, slq_num_matvecs=k, probes=probes))
    torch.cuda.synchronize(); dt = time.perf_counter() - t0
    print(f"M={M} |X|={batch} st_samples={st_samples} slq k={k} x2 probes: objective={val:.6g}  {dt:.3f} s  launches={L.lip_launch_count()-l0}", flush=True)
