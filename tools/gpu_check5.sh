set -x
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -3 gpurun_out/pytest_gpu.log
python bench.py --no-cpu --no-slq > gpurun_out/bench_overlap.json 2> gpurun_out/bench_overlap.err; cat gpurun_out/bench_overlap.json; tail -3 gpurun_out/bench_overlap.err
LIP_SPLIT_OVERLAP=0 python bench.py --no-cpu --no-slq --no-e2e > gpurun_out/bench_nooverlap.json 2> gpurun_out/bench_nooverlap.err; cat gpurun_out/bench_nooverlap.json
