set -x
mkdir -p gpurun_out
python bench.py --workload lenet5 --no-cpu --steps 5 --warmup 2 > gpurun_out/bench_lenet5.json 2> gpurun_out/bench_lenet5.err; cat gpurun_out/bench_lenet5.json; tail -5 gpurun_out/bench_lenet5.err
python bench.py --workload lenet5 --no-cpu --no-e2e --steps 1 --warmup 1 > gpurun_out/plain_lenet.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/launches_lenet5.csv python bench.py --workload lenet5 --no-cpu --no-e2e --steps 1 --warmup 1 > gpurun_out/ncu_lenet.log 2>&1
