#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -k "not config_parity" 2>&1 | tail -25 > gpurun_out/r2_tests_a.log
tail -8 gpurun_out/r2_tests_a.log
timeout 900 python -m pytest tests/test_gpu_config_parity.py -m gpu -q -s > gpurun_out/r2_tests_b.log 2>&1
grep -n "^C3a\|^\.C3a\|^FC3a\|C2 CG\|C3b\|ResNet\|passed\|failed\|^E  " gpurun_out/r2_tests_b.log | grep -v "print(" | head -40
timeout 300 python tools/slq_time.py 1 64,409 native > gpurun_out/r2_slq1.log 2>&1
timeout 300 python tools/slq_time.py 4 409 native >> gpurun_out/r2_slq1.log 2>&1
cat gpurun_out/r2_slq1.log
