#!/bin/bash
mkdir -p gpurun_out
timeout 300 python bench.py --workload lenet5 --no-cpu --no-e2e --no-extra --steps 1 --warmup 1 > gpurun_out/r2_lenet_plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:conv5 -s 4 -c 4 -o gpurun_out/r2_lenet_conv5 -f python bench.py --workload lenet5 --no-cpu --no-e2e --no-extra --steps 1 --warmup 1 > gpurun_out/r2_lenet_ncu.log 2>&1
ls -la gpurun_out/*.ncu-rep
