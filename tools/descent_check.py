"""Manual GPU check: a small step along -dZ from variational_grad_scalable lowers the (same-probes) scalable objective."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch
import bench
from lip_b200 import train_inducing as TI
M = int(sys.argv[1]) if len(sys.argv) > 1 else 50
bench.M_POINTS = M
ost, lst, Z = bench.build_states()
D = ost.flat()[0].size
dev = torch.device("cuda")
Zt = torch.as_tensor(Z, device=dev)
X = torch.rand(256, 784, device=dev, generator=torch.Generator(device=dev).manual_seed(1))
probes = torch.randint(0, 2, (64, D), device=dev, generator=torch.Generator(device=dev).manual_seed(2)).float() * 2 - 1
kw = dict(full_set_size=bench.N_FULL, slq_samples=2, slq_num_matvecs=int(0.8 * M), probes=probes)
loss0, g = TI.variational_grad_scalable(Zt, X, lst, bench.ALPHA, "classifier", 0, **kw)
print(f"M={M}: loss {float(loss0):.6e}  |g|={float(g.norm()):.4e}")
for step in (1e-4, 1e-3, 1e-2, 1e-1):
    eta = step / float(g.norm())
    lm = float(TI.alternative_objective_scalable(Zt - eta * g, X, lst, bench.ALPHA, "classifier", 0, **kw))
    lp = float(TI.alternative_objective_scalable(Zt + eta * g, X, lst, bench.ALPHA, "classifier", 0, **kw))
    pred = -eta * float((g * g).sum())
    print(f"  |dZ|={step:.0e}: loss(Z - eta g) - loss = {lm - float(loss0):+.4e}   loss(Z + eta g) - loss = {lp - float(loss0):+.4e}   "
          f"first-order prediction {pred:+.4e}")
# the same step measured on the deterministic exact-Gram objective (float64 value; differs from the scalable one by a Z-independent constant)
def exact(Zv):
    return float(TI._exact_value(TI._exact_parts(Zv, X, lst, bench.ALPHA, "classifier", bench.N_FULL))[0])
e0 = exact(Zt)
print(f"exact-Gram objective {e0:.9e}")
for step in (1e-3, 1e-2, 1e-1):
    eta = step / float(g.norm())
    em, ep = exact(Zt - eta * g), exact(Zt + eta * g)
    print(f"  |dZ|={step:.0e}: exact(Z - eta g) - exact = {em - e0:+.4e}   exact(Z + eta g) - exact = {ep - e0:+.4e}   "
          f"first-order prediction {-eta * float((g * g).sum()):+.4e}")
