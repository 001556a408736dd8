#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tools/dist_slq_check.py > gpurun_out/r2_dist_slq_2gpu.log 2>&1
echo rc=$?
grep -v "^W\|OMP_NUM\|^\*\*\*" gpurun_out/r2_dist_slq_2gpu.log | tail -30
