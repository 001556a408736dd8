"""Manual GPU harness: the small BASELINE configs that bench.py does not time.
  C1 toy sine regressor (D=241, M=40): curvature products/s at B=256
  C2 XOR classifier (D=354, M=32): Hutchinson trace with 1000 Rademacher probes + CG posterior solves for 1000 Gaussian RHS
  C3a subset-89-shaped classifier (D=2274, M=512): SLQ logdet k=409, 4 probes
  C3b MNIST-MLP: posterior sampler A^{-1/2} eps (Lanczos 2M on the Gram) + predictive at 256 test points, S=16
These are launch / latency bound (D <= 2274) or small dense problems: report wall time and rates only."""
import math, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch
import lip_b200
from lip_b200 import ggn, lla, matfree, stochtrace, sample as S, _cabi
from helpers import make_pair
L = _cabi.lib()
cu = lambda x: torch.as_tensor(np.asarray(x), dtype=torch.float32, device="cuda")

def timed(fn, reps=5):
    fn(); torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps): out = fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / reps, out

rng = np.random.default_rng(0)
# ---- C1
ost, lst = make_pair("regressor", hidden=[8, 8, 8, 8], n_out=1, in_dim=1, seed=1001)
Z = cu(rng.standard_normal((40, 1))); D = ost.flat()[0].size
cvp = lla.compute_curvature_approx(lst, Z, "regressor", 0.5, full_set_size=240)
V = cu(rng.choice([-1.0, 1.0], size=(256, D)))
dt, _ = timed(lambda: cvp(V), 50)
print(f"C1 toy-sine regressor D={D} M=40: curvature_vp B=256 in {dt*1e6:.0f} us -> {256/dt:,.0f} products/s", flush=True)
# ---- C2
ost, lst = make_pair("classifier", hidden=[16, 16], n_out=2, in_dim=2, seed=1002)
Z = cu(rng.standard_normal((32, 2))); D = ost.flat()[0].size
cvp = lla.compute_curvature_approx(lst, Z, "classifier", 9e-4, full_set_size=800)
eps = cu(rng.choice([-1.0, 1.0], size=(1000, D)))
dt, tr = timed(lambda: stochtrace.stochastic_trace_estimator_mvp(cvp, D, 0, eps=eps), 20)
print(f"C2 XOR classifier D={D} M=32: Hutchinson trace, 1000 probes: {dt*1e3:.2f} ms ({1000/dt:,.0f} products/s), estimate {float(tr):.5g}", flush=True)
rhs = cu(rng.standard_normal((1000, D)))
dt, (x, it) = timed(lambda: matfree.cg(cvp, rhs), 3)
res = (cvp(x) - rhs).norm(dim=1) / rhs.norm(dim=1)
print(f"C2 CG posterior solves, 1000 Gaussian RHS, tol 1e-5: {dt*1e3:.1f} ms, iterations {int(it.min())}..{int(it.max())}, max rel residual {float(res.max()):.2e}", flush=True)
# ---- C3a
ost, lst = make_pair("classifier", hidden=[32, 32, 32], n_out=2, in_dim=2, seed=1003)
Z = cu(rng.standard_normal((512, 2))); D = ost.flat()[0].size
Wz, WzT = ggn.compute_W_vps(lst, Z, "classifier", full_set_size=None)
sa = math.sqrt(0.5); d = 512 * 2
Av = matfree.batched(lambda v: torch.cat([sa * v.reshape(-1, D), WzT(v.reshape(-1, D)).reshape(-1, d)], dim=1))
vA = matfree.batched(lambda u: Wz(u.reshape(-1, D + d)[:, D:].reshape(-1, 512, 2)).add_(u.reshape(-1, D + d)[:, :D], alpha=sa))
problem = matfree.funm.integrand_funm_product_logdet(matfree.decomp.bidiag(409))
pr = cu(rng.choice([-1.0, 1.0], size=(4, D)))
dt, val = timed(lambda: problem(Av, pr, vA).mean(), 1)
print(f"C3a subset-89-shaped classifier D={D} M=512: SLQ logdet (GKL k=409, 4 probes) {dt:.3f} s, estimate {float(val):.5g}", flush=True)
# ---- C3b sampler + predictive
import bench
ost, lst, Zh = bench.build_states()
D = ost.flat()[0].size
Zd = cu(Zh)
Sn = 16
Eps = torch.randn(Sn, D, device="cuda")
t0 = time.perf_counter()
w = S.sample(lst, Zd, D, 1e-3, 0, "classifier", num_samples=Sn, full_set_size=60000, eps=Eps)
torch.cuda.synchronize(); t1 = time.perf_counter()
Xnew = cu(rng.random((256, 784)))
pred = lla.predict_lla_scalable(lst, Xnew, Zd, "classifier", 1e-3, full_set_size=60000, num_samples=Sn, eps=Eps)
torch.cuda.synchronize(); t2 = time.perf_counter()
print(f"C3b MNIST-MLP D={D} M=512 (d=5120): sample() S={Sn} (Gram + LU + Lanczos-1024 + W/W^T) {t1-t0:.2f} s; "
      f"predict_lla_scalable(256 test points, S={Sn}) incl. its own sample() {t2-t1:.2f} s; finite={bool(torch.isfinite(pred).all())} "
      f"pred std over samples {float(pred.std(0).mean()):.3g}", flush=True)
