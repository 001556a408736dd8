#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_config_parity.py -m gpu -q -s -k "C3a or C2 or C3b_slq" > gpurun_out/r2_tests_b.log 2>&1
grep -n "C3a\|C2 CG\|C3b\|passed\|failed" gpurun_out/r2_tests_b.log | grep -v "^.*print(" | head -40
