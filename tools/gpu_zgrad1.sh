mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_zgrad.py -m gpu -q > gpurun_out/pytest_zgrad.log 2>&1
echo "pytest rc=$?"; tail -40 gpurun_out/pytest_zgrad.log
