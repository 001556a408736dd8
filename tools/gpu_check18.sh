set -x
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -8 gpurun_out/pytest_gpu.log
python tools/objective_time.py 50 256 64 40 > gpurun_out/objective_time.log 2>&1
python tools/objective_time.py 512 256 64 409 >> gpurun_out/objective_time.log 2>&1
cat gpurun_out/objective_time.log
