#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -8 > gpurun_out/r2_tests_all.log
tail -5 gpurun_out/r2_tests_all.log
timeout 800 python tools/resnet_step_time.py > gpurun_out/r2_resnet_step.log 2>&1; tail -8 gpurun_out/r2_resnet_step.log
