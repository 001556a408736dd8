#!/bin/bash
mkdir -p gpurun_out
cat > /tmp/slq_b1.py <<'PY'
import os, sys
ROOT = os.environ.get("GRAFT_REPO_ROOT", "/root/repo")
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch, bench
from lip_b200 import ggn, matfree
ost, lst, Z = bench.build_states()
D = ost.flat()[0].size
Zd = torch.as_tensor(Z, device="cuda")
Wz, WzT = ggn.compute_W_vps(lst, Zd, "classifier", full_set_size=None)
Av = matfree.gkl_target(WzT, Wz, bench.ALPHA)
P = (torch.randint(0, 2, (1, D), device="cuda").float() * 2 - 1)
q = matfree.slq_quadrature(Av, P, 409, form="gkl")
torch.cuda.synchronize()
print(q)
PY
timeout 300 python /tmp/slq_b1.py > gpurun_out/r2_b1_plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -s 12000 -c 120 --csv --log-file gpurun_out/r2_launches_slq_b1_reduced_step300.csv python /tmp/slq_b1.py > gpurun_out/r2_b1_ncu.log 2>&1
python tools/summarize_launches.py gpurun_out/r2_launches_slq_b1_step300.csv 25
