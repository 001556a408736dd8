"""Manual GPU harness: relative error of the Z-gradient of the KL objective against the float64 autograd oracle on a toy classifier as
alpha shrinks and beta = N/M grows (the scale configs sit at beta/alpha ~ 1e6): Woodbury-space form vs Gram-space form."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch
from helpers import make_pair, rel_err
from oracle import lip_oracle as O
from lip_b200 import train_inducing as TI
cu = lambda a: torch.as_tensor(np.asarray(a), dtype=torch.float32, device="cuda")
ost, lst = make_pair("classifier", hidden=[16, 16], n_out=3, in_dim=2, seed=31)
rng = np.random.default_rng(32)
Z = rng.standard_normal((12, 2)).astype(np.float32); X = rng.standard_normal((40, 2)).astype(np.float32)
D = ost.flat()[0].size
probes = rng.choice([-1.0, 1.0], size=(64, D)).astype(np.float32)
for alpha, N in ((0.05, 800), (1e-2, 6000), (1e-3, 6000), (1e-3, 60000)):
    _, ref_g = O.variational_grad_dense(Z, X, ost, alpha, "classifier", full_set_size=N)
    _, g = TI.variational_grad_scalable(cu(Z), cu(X), lst, alpha, "classifier", 0, full_set_size=N, slq_num_matvecs=4, probes=cu(probes), gradient="woodbury")
    _, g2 = TI.variational_grad_scalable_exact(cu(Z), cu(X), lst, alpha, "classifier", 0, full_set_size=N)
    _, g3 = TI.variational_grad_scalable_exact(cu(Z), cu(X), lst, alpha, "classifier", 0, full_set_size=N, accurate_grams=True)
    g = g.cpu().numpy(); g2 = g2.cpu().numpy(); g3 = g3.cpu().numpy()
    print(f"alpha={alpha} N={N} beta={N/12:.0f}: |g_ref|={np.linalg.norm(ref_g):.3e}  woodbury-form rel err {rel_err(g, ref_g):.2e}  gram-form rel err {rel_err(g2, ref_g):.2e}  gram-form with f64 Grams of the fp32 factors {rel_err(g3, ref_g):.2e}")
