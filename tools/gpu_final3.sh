set -x
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -3 gpurun_out/pytest_gpu.log
python bench.py --workload lenet5 --no-cpu --steps 5 --warmup 2 > gpurun_out/bench_lenet5.json 2> gpurun_out/bench_lenet5.err
python bench.py --workload resnet1m --probes 16 --no-cpu --no-e2e --steps 3 --warmup 1 > gpurun_out/bench_resnet1m.json 2> gpurun_out/bench_resnet1m.err
python bench.py > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err
for f in bench_lenet5 bench_resnet1m bench_default; do cut -c1-170 gpurun_out/$f.json; done
