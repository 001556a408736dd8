"""Manual GPU harness: time lip_zgrad (GGN mode, ubar = v: gradient of a probe block's quadratic forms) on the
headline MNIST-MLP shape, and one inducing-point training step (train_inducing.optimize_step, scalable objective + Hutchinson
Z-gradient) with the reference config's sizes (config/scale/mlp_mnist.yml)."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch
import bench
from lip_b200 import train_inducing, lla, utils, _cabi
M = int(sys.argv[1]) if len(sys.argv) > 1 else 512
B = int(sys.argv[2]) if len(sys.argv) > 2 else 64
batch = int(sys.argv[3]) if len(sys.argv) > 3 else 256
k = int(sys.argv[4]) if len(sys.argv) > 4 else int(0.8 * M)
bench.M_POINTS = M
ost, lst, Z = bench.build_states()
D = ost.flat()[0].size
dev = torch.device("cuda")
Zt = torch.as_tensor(Z, device=dev)
probes = torch.randint(0, 2, (B, D), device=dev).float() * 2 - 1
L = _cabi.lib()
cvp = lla.compute_curvature_approx(lst, Zt, "classifier", bench.ALPHA, full_set_size=bench.N_FULL)
sigma = [4 * a * b for a, b in zip([784, 1024, 512, 256, 128], [1024, 512, 256, 128, 10])]
# algorithmic FLOP per (probe, point): forward tangent (2 GEMMs / layer), reverse (3 GEMMs / layer), x2 sides (ubar and v)
s_all = sum(a * b for a, b in zip([784, 1024, 512, 256, 128], [1024, 512, 256, 128, 10]))
s_ge2 = s_all - 784 * 1024
flop = 2 * B * M * 2 * ((s_all + s_ge2) + (2 * s_all + s_ge2))
for rep in range(3):
    torch.cuda.synchronize(); l0 = L.lip_launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); g = cvp.zgrad(probes, probes); e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    print(f"lip_zgrad GGN mode: M={M} B={B} D={D}: {ms:.2f} ms  ({B / ms * 1e3:.0f} probe-gradients/s, {flop / ms / 1e9:.1f} TFLOP/s algorithmic, {cvp._lip_model.path_name()}) "
          f"launches={L.lip_launch_count() - l0} |dZ|={float(g.norm()):.4g}", flush=True)
rng = np.random.default_rng(9)
X = torch.as_tensor(rng.random((batch, 784), dtype=np.float32), device=dev)
opt = utils.adam(1e-3)
state = opt.init(Zt)
Zc = Zt
for rep in range(3):
    torch.cuda.synchronize(); l0 = L.lip_launch_count(); t0 = time.perf_counter()
    Zc, state, loss = train_inducing.optimize_step(Zc, X, lst, bench.ALPHA, state, rep, opt, None, "classifier", full_set_size=bench.N_FULL,
                                                   scalable=True, st_samples=B, slq_samples=2, slq_num_matvecs=k)
    torch.cuda.synchronize(); dt = time.perf_counter() - t0
    print(f"optimize_step (scalable objective + exact dZ): M={M} |X|={batch} st_samples={B} slq k={k}: loss={float(loss):.6g} {dt:.3f} s "
          f"launches={L.lip_launch_count() - l0} |dZ step|={float((Zc - Zt).norm()):.4g}", flush=True)
