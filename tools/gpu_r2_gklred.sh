#!/bin/bash
mkdir -p gpurun_out
LIP_GKL_DEBUG=1 timeout 1200 python -m pytest tests/test_gpu_config_parity.py tests/test_gpu_parity.py -m gpu -q -s -k "slq or gkl or logdet or C3a or C3b or config" 2>&1 | grep -E "passed|failed|lip gkl|FAILED|rel:" | tail -20
for r in 1 0; do
echo "== LIP_GKL_REDUCED=$r"
LIP_GKL_DEBUG=1 LIP_GKL_REDUCED=$r timeout 600 python tools/slq_native_time.py 4 409 2>&1 | tail -3
LIP_GKL_DEBUG=1 LIP_GKL_REDUCED=$r timeout 600 python tools/slq_native_time.py 1 409 2>&1 | tail -3
done
