"""Manual GPU harness: ONE native lip_slq_quadrature call (GKL form, headline C3b operator): time and value.
usage: python tools/slq_native_time.py [probes] [k]     (LIP_GKL_REDUCED=0/1 selects the explicit / reduced u basis)"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import torch
import bench
from lip_b200 import ggn, matfree
ost, lst, Z = bench.build_states()
D = ost.flat()[0].size
dev = torch.device("cuda")
Zd = torch.as_tensor(Z, device=dev)
Wz, WzT = ggn.compute_W_vps(lst, Zd, "classifier", full_set_size=None)
Av = matfree.gkl_target(WzT, Wz, bench.ALPHA)
ns = int(sys.argv[1]) if len(sys.argv) > 1 else 4
k = int(sys.argv[2]) if len(sys.argv) > 2 else 409
g = torch.Generator(device=dev); g.manual_seed(5)
probes = (torch.randint(0, 2, (ns, D), device=dev, generator=g).float() * 2 - 1)
for rep in range(2):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    q = matfree.slq_quadrature(Av, probes, k, form="gkl", fn="log", clip_min=None)
    torch.cuda.synchronize(); dt = time.perf_counter() - t0
print(f"native GKL SLQ k={k} probes={ns} reduced={os.environ.get('LIP_GKL_REDUCED', '1')}: {dt:.3f} s  quadratures {q.cpu().numpy()}  mean {q.double().mean().item():.8g}")
