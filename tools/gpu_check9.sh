set -x
mkdir -p gpurun_out
python bench.py --no-cpu --no-slq --no-e2e > gpurun_out/bench_wide.json 2> gpurun_out/bench_wide.err; cut -c1-200 gpurun_out/bench_wide.json; tail -3 gpurun_out/bench_wide.err
python bench.py --no-cpu --no-slq --no-e2e --steps 2 --warmup 1 > gpurun_out/plain_w.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/launches_wide.csv python bench.py --no-cpu --no-slq --no-e2e --steps 2 --warmup 1 > gpurun_out/ncu_w.log 2>&1
