set -x
mkdir -p gpurun_out
python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/smoke.log; tail -2 gpurun_out/smoke.log
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -3 gpurun_out/pytest_gpu.log
python bench.py > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err
python bench.py --workload resnet1m --probes 16 --no-cpu --no-e2e --steps 5 --warmup 3 > gpurun_out/bench_resnet1m.json 2> gpurun_out/bench_resnet1m.err
python bench.py --workload lenet5 --no-cpu --steps 5 --warmup 3 > gpurun_out/bench_lenet5.json 2> gpurun_out/bench_lenet5.err
for f in bench_lenet5 bench_resnet1m bench_default; do cut -c1-170 gpurun_out/$f.json; done
