#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q -k "not config_parity" 2>&1 | tail -25 > gpurun_out/r2_tests_a.log
tail -5 gpurun_out/r2_tests_a.log
timeout 900 python -m pytest tests/test_gpu_config_parity.py -m gpu -q -s > gpurun_out/r2_tests_b.log 2>&1
grep -n "^C3a\|^\.C3a\|^FC3a\|C2 CG\|C3b\|ResNet\|passed\|failed\|^E  " gpurun_out/r2_tests_b.log | grep -v "print(" | head -60
timeout 300 python tools/slq_time.py 1 64 native > gpurun_out/r2_slq1.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file gpurun_out/r2_launches_slq1.csv python tools/slq_time.py 1 64 native > gpurun_out/r2_ncu.log 2>&1
tail -3 gpurun_out/r2_slq1.log; tail -3 gpurun_out/r2_ncu.log
