#!/bin/bash
mkdir -p gpurun_out
A="--no-cpu --no-slq --no-e2e --no-extra --no-train-step --steps 1 --warmup 1"
timeout 300 python bench.py $A > gpurun_out/r2_traffic_plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_traffic_step.csv python bench.py $A > gpurun_out/r2_traffic_ncu.log 2>&1
python tools/traffic_summary.py gpurun_out/r2_traffic_step.csv gpurun_out/r02_traffic_ggn_vp.json
