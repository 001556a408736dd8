"""debug: C3a GKL k=409 — where do NaNs appear?"""
import os, sys, importlib.util
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch
import lip_b200
from lip_b200 import ggn, matfree, toymodels, scalemodels
spec = importlib.util.spec_from_file_location("g", os.path.join(ROOT, "tests", "golden", "make_golden_configs.py")); g = importlib.util.module_from_spec(spec); spec.loader.exec_module(g)
ost, Z, cfg, probes, eps = g.c3a_inputs()
lst = scalemodels.TrainState(params=ost.params, apply_fn=toymodels.SimpleClassifier(32, 3, 2).apply, batch_stats=ost.batch_stats)
cu = lambda x: torch.as_tensor(np.asarray(x), dtype=torch.float32, device="cuda")
Wz, WzT = ggn.compute_W_vps(lst, cu(Z), "classifier", full_set_size=None)
Av = matfree.gkl_target(WzT, Wz, cfg["alpha"])
for rep in range(6):
    res, _ = matfree.decomp.bidiag(409)(Av, Av._lip_transpose, cu(probes))
    al, be = res.alphas.cpu().numpy(), res.betas.cpu().numpy()
    for b in range(4):
        bad_a = np.where(~np.isfinite(al[b]))[0]; bad_b = np.where(~np.isfinite(be[b]))[0]
        i = min(list(bad_a[:1]) + list(bad_b[:1]) + [409])
        lo = max(0, i - 3)
        print(f"rep {rep} probe {b}: first bad alpha {bad_a[:1]} beta {bad_b[:1]} min|beta[1:]| {np.nanmin(np.abs(be[b][1:])):.3e} at {np.nanargmin(np.abs(be[b][1:]))+1}; around: alphas {al[b][lo:i+2]} betas {be[b][lo:i+2]}")
