#!/bin/bash
mkdir -p gpurun_out
timeout 300 python bench.py --no-slq --no-extra --no-train-step --no-cpu --no-e2e --steps 2 --warmup 1 > gpurun_out/r2_head_plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 500 --csv --log-file gpurun_out/r2_launches_bench_headfused.csv python bench.py --no-slq --no-extra --no-train-step --no-cpu --no-e2e --steps 2 --warmup 1 > gpurun_out/r2_head_ncu.log 2>&1
python tools/summarize_launches.py gpurun_out/r2_launches_bench_headfused.csv 18
