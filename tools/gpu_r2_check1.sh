#!/bin/bash
# round 2, first GPU check: the native Krylov entry points under the whole GPU suite + config-size parity + SLQ timing
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q -k "not config_parity" 2>&1 | tail -15 > gpurun_out/r2_tests_a.log
cat gpurun_out/r2_tests_a.log
timeout 900 python -m pytest tests/test_gpu_config_parity.py -m gpu -q -s 2>&1 | tail -60 > gpurun_out/r2_tests_b.log
cat gpurun_out/r2_tests_b.log
timeout 300 python tools/slq_time.py 4 64,409 native 2>&1 | tail -8 > gpurun_out/r2_slq_native.log
timeout 300 python tools/slq_time.py 1 409 native 2>&1 | tail -4 >> gpurun_out/r2_slq_native.log
timeout 300 python tools/slq_time.py 4 64 callback 2>&1 | tail -4 >> gpurun_out/r2_slq_native.log
cat gpurun_out/r2_slq_native.log
