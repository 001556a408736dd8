set -x
mkdir -p gpurun_out
python bench.py --workload resnet1m --probes 16 --no-cpu --no-e2e --steps 3 --warmup 1 > gpurun_out/bench_resnet1m.json 2> gpurun_out/bench_resnet1m.err
timeout 600 python bench.py --workload resnet1m --points 4096 --probes 4 --no-cpu --no-e2e --steps 1 --warmup 1 > gpurun_out/bench_resnet1m_m4096.json 2> gpurun_out/bench_resnet1m_m4096.err
python bench.py --workload lenet5 --no-cpu --steps 5 --warmup 2 > gpurun_out/bench_lenet5.json 2> gpurun_out/bench_lenet5.err
python bench.py > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err
python bench.py --workload resnet1m --probes 16 --no-cpu --no-e2e --steps 1 --warmup 1 > gpurun_out/plain_resnet.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 1300 --csv --log-file gpurun_out/launches_resnet1m.csv python bench.py --workload resnet1m --probes 16 --no-cpu --no-e2e --steps 1 --warmup 1 > gpurun_out/ncu_resnet.log 2>&1
for f in bench_resnet1m bench_resnet1m_m4096 bench_lenet5 bench_default; do cut -c1-170 gpurun_out/$f.json; done
