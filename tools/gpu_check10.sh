set -x
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -3 gpurun_out/pytest_gpu.log
python tools/tc_accuracy_exp.py > gpurun_out/tc_accuracy_wide.log 2>&1; cat gpurun_out/tc_accuracy_wide.log
python bench.py --no-cpu --no-slq --no-e2e > gpurun_out/bench_w10.json 2> gpurun_out/bench_w10.err; cut -c1-200 gpurun_out/bench_w10.json
LIP_TC_KC=16 python bench.py --no-cpu --no-slq --no-e2e > gpurun_out/bench_w10_kc16.json 2> gpurun_out/bench_w10_kc16.err; cut -c1-200 gpurun_out/bench_w10_kc16.json
