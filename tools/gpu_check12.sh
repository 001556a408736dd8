set -x
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -15 gpurun_out/pytest_gpu.log
python bench.py --no-cpu > gpurun_out/bench_splitm.json 2> gpurun_out/bench_splitm.err; cat gpurun_out/bench_splitm.json; tail -3 gpurun_out/bench_splitm.err
