mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_golden_fixtures.py tests/test_gpu_zgrad.py tests/test_gpu_objectives.py -m gpu -q > gpurun_out/pytest_f1.log 2>&1
echo "pytest rc=$?"; tail -5 gpurun_out/pytest_f1.log
timeout 600 python tools/zgrad_time.py 512 64 256 409 > gpurun_out/zgrad_time.txt 2>&1; cat gpurun_out/zgrad_time.txt | tail -8
timeout 600 python tools/zgrad_time.py 50 64 256 40 >> gpurun_out/zgrad_time.txt 2>&1; tail -6 gpurun_out/zgrad_time.txt
