"""Manual GPU harness: where one inducing-point training step (train_inducing.optimize_step, scalable objective) spends its time."""
import os, sys, time, math
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch
import bench
from lip_b200 import train_inducing as TI, lla, matfree, stochtrace, _cabi
M = int(sys.argv[1]) if len(sys.argv) > 1 else 512
B = int(sys.argv[2]) if len(sys.argv) > 2 else 64
k = int(sys.argv[3]) if len(sys.argv) > 3 else int(0.8 * M)
bench.M_POINTS = M
ost, lst, Z = bench.build_states()
D = ost.flat()[0].size
dev = torch.device("cuda")
Zt = torch.as_tensor(Z, device=dev)
X = torch.rand(256, 784, device=dev)
probes = torch.randint(0, 2, (B, D), device=dev).float() * 2 - 1
alpha, N = bench.ALPHA, bench.N_FULL
def timed(name, fn, reps=2):
    out = None
    for _ in range(reps):
        torch.cuda.synchronize(); t0 = time.perf_counter(); out = fn(); torch.cuda.synchronize(); dt = time.perf_counter() - t0
    print(f"  {name:48s} {dt * 1e3:9.1f} ms", flush=True)
    return out
print(f"M={M} |X|=256 probes={B} slq k={k} x 2")
parts = timed("parts: bind X and Z, Gram W_z^T W_z, LU", lambda: TI._scalable_parts(Zt, X, lst, alpha, "classifier", N))
S_vp, Wz, WzT, Sz_inv = parts["S_vp"], parts["Wz"], parts["WzT"], parts["Sz_inv"]
comp = matfree.batched(lambda V: S_vp(Sz_inv(V)))
timed("hutchpp_v2 trace (s1 = B - 16, s2 = 16)", lambda: stochtrace.hutchpp_v2(comp, lambda _: probes, s1=B - 16, s2=16))
timed("whole objective (shared parts)", lambda: TI.alternative_objective_scalable(Zt, X, lst, alpha, "classifier", 0, full_set_size=N, slq_samples=2, slq_num_matvecs=k, probes=probes, _parts=parts))
Sz_vp = lla.compute_curvature_approx(lst, Zt, "classifier", alpha, full_set_size=N)
def grad():
    a = Sz_inv(probes); b = probes - Sz_inv(S_vp(probes)); return Sz_vp.zgrad(a, b)
timed("Hutchinson dZ (2 S_Z^-1, 1 S_X, lip_zgrad)", grad)
timed("variational_grad_scalable (everything)", lambda: TI.variational_grad_scalable(Zt, X, lst, alpha, "classifier", 0, full_set_size=N, slq_samples=2, slq_num_matvecs=k, probes=probes))
