#!/bin/bash
mkdir -p gpurun_out
timeout 300 python tools/slq_time.py 4 200 native > gpurun_out/r2_ncu_slq_plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:"project_kernel|subtract_kernel" -s 776 -c 4 -o gpurun_out/r2_prof_slq python tools/slq_time.py 4 200 native > gpurun_out/r2_ncu_slq.log 2>&1
tail -3 gpurun_out/r2_ncu_slq_plain.log; tail -3 gpurun_out/r2_ncu_slq.log; ls -la gpurun_out/*.ncu-rep
