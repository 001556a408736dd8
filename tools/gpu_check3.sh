set -x
mkdir -p gpurun_out
python tools/tc_selftest.py > gpurun_out/selftest.log 2>&1; tail -2 gpurun_out/selftest.log
LIP_TC_2CTA=1 python tools/tc_selftest.py > gpurun_out/selftest_2cta.log 2>&1; tail -2 gpurun_out/selftest_2cta.log
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -3 gpurun_out/pytest_gpu.log
python tools/tc_microbench.py 20 > gpurun_out/microbench.log 2>&1
cat gpurun_out/microbench.log
python bench.py --no-cpu --no-slq > gpurun_out/bench_1cta.json 2> gpurun_out/bench_1cta.err; cat gpurun_out/bench_1cta.json
LIP_TC_2CTA=1 python bench.py --no-cpu --no-slq > gpurun_out/bench_2cta.json 2> gpurun_out/bench_2cta.err; cat gpurun_out/bench_2cta.json
