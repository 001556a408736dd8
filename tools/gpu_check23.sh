mkdir -p gpurun_out
timeout 500 python tools/tc_small_m.py > gpurun_out/tc_small_m.txt 2>&1; tail -5 gpurun_out/tc_small_m.txt
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -3 gpurun_out/pytest_gpu.log
timeout 600 python tools/zgrad_time.py 50 64 256 40 > gpurun_out/zgrad_time_m50.txt 2>&1; tail -3 gpurun_out/zgrad_time_m50.txt
timeout 600 python tools/objective_time.py 50 256 64 40 > gpurun_out/objective_time_m50.txt 2>&1; tail -2 gpurun_out/objective_time_m50.txt
