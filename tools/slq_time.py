"""Manual GPU harness: SLQ logdet (GKL and Lanczos forms) timing vs k on the headline config, with the implied re-orth bandwidth.
usage: python tools/slq_time.py [probes] [k,k,...] [native|callback]"""
import math, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np, torch
import bench
from lip_b200 import ggn, lla, matfree, matfree_monkeypatch, _cabi
ost, lst, Z = bench.build_states()
D = ost.flat()[0].size
dev = torch.device("cuda")
Zd = torch.as_tensor(Z, device=dev)
Wz, WzT = ggn.compute_W_vps(lst, Zd, "classifier", full_set_size=None)
sa = math.sqrt(bench.ALPHA); d = bench.M_POINTS * bench.DIMS[-1]; M = bench.M_POINTS; K = bench.DIMS[-1]
mode = sys.argv[3] if len(sys.argv) > 3 else "native"
if mode == "native":
    Av = matfree.gkl_target(WzT, Wz, bench.ALPHA)
    vA = Av._lip_transpose
else:
    Av = matfree.batched(lambda v: torch.cat([sa * v.reshape(-1, D), WzT(v.reshape(-1, D)).reshape(-1, d)], dim=1))
    vA = matfree.batched(lambda u: Wz(u.reshape(-1, D + d)[:, D:].reshape(-1, M, K)).add_(u.reshape(-1, D + d)[:, :D], alpha=sa))
cvp = lla.compute_curvature_approx(lst, Zd, "classifier", bench.ALPHA, full_set_size=bench.N_FULL)
ns = int(sys.argv[1]) if len(sys.argv) > 1 else 4
probes = (torch.randint(0, 2, (ns, D), device=dev).float() * 2 - 1)
L = _cabi.lib()
for k in [int(x) for x in (sys.argv[2] if len(sys.argv) > 2 else "16,64,128,409").split(",")]:
    problem = matfree.funm.integrand_funm_product_logdet(matfree.decomp.bidiag(k))
    for rep in range(2):
        torch.cuda.synchronize(); l0 = L.lip_launch_count(); t0 = time.perf_counter()
        val = problem(Av, probes, vA).mean().item()
        torch.cuda.synchronize(); dt = time.perf_counter() - t0
    n1, n2 = D + d, D
    reorth_bytes = ns * 4 * sum(2 * i * n1 + 2 * (i + 1) * n2 for i in range(k))
    print(f"[{mode}] GKL k={k:4d} probes={ns}: {dt:8.3f} s  logdet~{val:.6g}  launches={L.lip_launch_count()-l0}  re-orth algorithmic bytes {reorth_bytes/1e9:8.1f} GB -> {reorth_bytes/dt/1e12:5.2f} TB/s if it were all of the time", flush=True)
    lz = matfree_monkeypatch.integrand_funm_sym_logdet(matfree.decomp.tridiag_sym(k))
    for rep in range(2):
        torch.cuda.synchronize(); l0 = L.lip_launch_count(); t0 = time.perf_counter()
        val = lz(cvp, probes).mean().item()
        torch.cuda.synchronize(); dt = time.perf_counter() - t0
    lz_bytes = ns * 4 * sum(2 * (i + 1) * D * 2 for i in range(k))
    print(f"[{mode}] Lanczos k={k:4d} probes={ns}: {dt:8.3f} s  logdet(clip1)~{val:.6g}  launches={L.lip_launch_count()-l0}  re-orth {lz_bytes/1e9:8.1f} GB -> {lz_bytes/dt/1e12:5.2f} TB/s", flush=True)
