set -x
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -3 gpurun_out/pytest_gpu.log
python tools/tc_accuracy_exp.py > gpurun_out/tc_accuracy_wide.log 2>&1; head -3 gpurun_out/tc_accuracy_wide.log
python bench.py --no-cpu > gpurun_out/bench_wide.json 2> gpurun_out/bench_wide.err; cat gpurun_out/bench_wide.json; tail -3 gpurun_out/bench_wide.err
LIP_TC_WIDE=0 python bench.py --no-cpu --no-slq --no-e2e > gpurun_out/bench_nowide.json 2> gpurun_out/bench_nowide.err
