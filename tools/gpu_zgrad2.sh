mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_objectives.py -m gpu -q > gpurun_out/pytest_obj.log 2>&1
echo "pytest rc=$?"; tail -60 gpurun_out/pytest_obj.log
