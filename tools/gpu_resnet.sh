set -x
mkdir -p gpurun_out
python bench.py --workload resnet1m --probes 16 --no-cpu --no-e2e --steps 2 --warmup 1 > gpurun_out/bench_resnet1m.json 2> gpurun_out/bench_resnet1m.err; cat gpurun_out/bench_resnet1m.json; tail -5 gpurun_out/bench_resnet1m.err
python bench.py --workload resnet1m --probes 16 --no-cpu --no-e2e --steps 1 --warmup 1 > gpurun_out/plain_resnet.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/launches_resnet1m.csv python bench.py --workload resnet1m --probes 16 --no-cpu --no-e2e --steps 1 --warmup 1 > gpurun_out/ncu_resnet.log 2>&1
