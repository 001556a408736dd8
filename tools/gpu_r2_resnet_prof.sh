#!/bin/bash
mkdir -p gpurun_out
A="--workload resnet1m --points 4096 --probes 4 --no-cpu --no-e2e --no-extra --no-train-step --steps 1 --warmup 1"
timeout 600 python bench.py $A > gpurun_out/r2_resnet_plain.log 2>gpurun_out/r2_resnet_plain.err && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/r2_launches_resnet1m_m4096.csv python bench.py $A > gpurun_out/r2_resnet_ncu.log 2>&1
python tools/summarize_launches.py gpurun_out/r2_launches_resnet1m_m4096.csv 30
tail -c 600 gpurun_out/r2_resnet_plain.log
