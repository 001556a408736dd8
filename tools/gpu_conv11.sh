mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "resnet or conv or lenet" > gpurun_out/pytest_resnet.log 2>&1
echo "pytest rc=$?"; tail -3 gpurun_out/pytest_resnet.log
python bench.py --workload resnet1m --probes 16 --no-cpu --no-e2e --steps 5 --warmup 2 > gpurun_out/bench_resnet1m.json 2> gpurun_out/bench_resnet1m.err; cut -c1-200 gpurun_out/bench_resnet1m.json; tail -5 gpurun_out/bench_resnet1m.err
python bench.py --workload resnet1m --probes 32 --no-cpu --no-e2e --steps 5 --warmup 2 > gpurun_out/bench_resnet1m_p32.json 2> gpurun_out/bench_resnet1m_p32.err; cut -c1-200 gpurun_out/bench_resnet1m_p32.json
ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file gpurun_out/launches_resnet1m.csv python bench.py --workload resnet1m --probes 16 --no-cpu --no-e2e --steps 1 --warmup 1 > gpurun_out/ncu_resnet.log 2>&1
