#!/bin/bash
N=${1:-2}
mkdir -p gpurun_out
if [ "$N" = "1" ]; then
  timeout 600 python tools/dist_slq_check.py > gpurun_out/r2_dist_slq_${N}gpu.log 2>&1
else
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 tools/dist_slq_check.py > gpurun_out/r2_dist_slq_${N}gpu.log 2>&1
fi
echo rc=$?
grep -v "^W1\|OMP_NUM\|^\*\*\*\|^\[rank" gpurun_out/r2_dist_slq_${N}gpu.log | tail -25
