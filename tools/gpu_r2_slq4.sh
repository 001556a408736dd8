#!/bin/bash
N=${1:-4}
mkdir -p gpurun_out
for r in 1 0 1; do
LIP_GKL_REDUCED=$r timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $N --steps 3 --warmup 3 --no-e2e --no-cpu --no-extra --no-train-step 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); s=d['slq_logdet']; print('reduced=$r N=$N slq', round(s['seconds'],4), s['layout']['probe_groups'], s['layout']['basis_shards_per_group'], s['launches_per_logdet'], 'lanczos', round(s['lanczos_form']['seconds'],4), s['logdet_estimate'])"
done
