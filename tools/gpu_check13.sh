set -x
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -5 gpurun_out/pytest_gpu.log
python bench.py > gpurun_out/bench_head.json 2> gpurun_out/bench_head.err; cat gpurun_out/bench_head.json; tail -3 gpurun_out/bench_head.err
python tools/slq_time.py 4 64,409 > gpurun_out/slq_time2.log 2>&1; cat gpurun_out/slq_time2.log
