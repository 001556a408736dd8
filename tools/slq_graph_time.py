"""Manual GPU harness: the k = 409 GKL logdet quadrature (lip_slq_quadrature) eager vs replayed from ONE CUDA graph."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch
import bench
from lip_b200 import ggn, lla, matfree
ost, lst, Z = bench.build_states()
D = ost.flat()[0].size
dev = torch.device("cuda")
Zd = torch.as_tensor(Z, device=dev)
Wz, WzT = ggn.compute_W_vps(lst, Zd, "classifier", full_set_size=None)
Av = matfree.gkl_target(WzT, Wz, bench.ALPHA)
k = int(sys.argv[1]) if len(sys.argv) > 1 else 409
for ns in (1, 4):
    P = (torch.randint(0, 2, (ns, D), device=dev).float() * 2 - 1)
    for rep in range(2):
        torch.cuda.synchronize(); t0 = time.perf_counter()
        q = matfree.slq_quadrature(Av, P, k, form="gkl")
        torch.cuda.synchronize(); dt = time.perf_counter() - t0
    print(f"eager  GKL k={k} probes={ns}: {dt:.3f} s  mean {q.mean().item():.8g}", flush=True)
    side = torch.cuda.Stream(); side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        matfree.slq_quadrature(Av, P, 4, form="gkl")
    torch.cuda.current_stream().wait_stream(side)
    g = torch.cuda.CUDAGraph()
    t0 = time.perf_counter()
    with torch.cuda.graph(g):
        out = matfree.slq_quadrature(Av, P, k, form="gkl")
    torch.cuda.synchronize(); tc = time.perf_counter() - t0
    for rep in range(3):
        torch.cuda.synchronize(); t0 = time.perf_counter()
        g.replay()
        torch.cuda.synchronize(); dt = time.perf_counter() - t0
    print(f"graph  GKL k={k} probes={ns}: {dt:.3f} s per replay (capture + instantiate {tc:.2f} s)  mean {out.mean().item():.8g}", flush=True)
    del g, out
