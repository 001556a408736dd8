#!/bin/bash
mkdir -p gpurun_out
timeout 300 python bench.py --workload lenet5 --no-cpu --no-e2e --no-extra --steps 2 --warmup 1 > gpurun_out/r2_lenet_plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 800 --csv --log-file gpurun_out/r2_launches_lenet5_fused.csv python bench.py --workload lenet5 --no-cpu --no-e2e --no-extra --steps 2 --warmup 1 > gpurun_out/r2_lenet_ncu.log 2>&1
python tools/summarize_launches.py gpurun_out/r2_launches_lenet5_fused.csv 25
