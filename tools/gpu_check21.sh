set -x
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -3 gpurun_out/pytest_gpu.log
python bench.py --workload resnet1m --probes 16 --no-cpu --no-e2e --steps 2 --warmup 1 > gpurun_out/bench_resnet1m.json 2> gpurun_out/bench_resnet1m.err; cut -c1-250 gpurun_out/bench_resnet1m.json; tail -3 gpurun_out/bench_resnet1m.err
