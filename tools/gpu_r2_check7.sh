#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_config_parity.py tests/test_gpu_parity.py -m gpu -q -k "C3a or C3b_slq or lanczos or G2 or sampler" 2>&1 | tail -4
timeout 300 python tools/slq_time.py 1 409 native 2>&1 | tail -2
timeout 300 python tools/slq_time.py 4 409 native 2>&1 | tail -2
