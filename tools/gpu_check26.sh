mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_objectives.py -m gpu -q 2>&1 | grep -v "Warning\|warnings.warn" | tail -4
python tools/descent_check.py 50 2>&1 | tail -6
python tools/zgrad_time.py 50 64 256 40 2>&1 | grep optimize_step
python tools/zgrad_time.py 512 64 256 409 2>&1 | grep optimize_step
python tools/lenet_step_time.py 2>&1 | grep optimize_step
