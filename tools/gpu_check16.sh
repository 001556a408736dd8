set -x
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -5 gpurun_out/pytest_gpu.log
python bench.py --workload lenet5 --no-cpu --steps 5 --warmup 2 > gpurun_out/bench_lenet5.json 2> gpurun_out/bench_lenet5.err; cut -c1-250 gpurun_out/bench_lenet5.json; tail -3 gpurun_out/bench_lenet5.err
timeout 600 python bench.py --workload resnet1m --points 4096 --probes 4 --no-cpu --no-e2e --steps 1 --warmup 1 > gpurun_out/bench_resnet1m_m4096.json 2> gpurun_out/bench_resnet1m_m4096.err; cut -c1-400 gpurun_out/bench_resnet1m_m4096.json; tail -3 gpurun_out/bench_resnet1m_m4096.err
nvidia-smi --query-gpu=memory.used --format=csv
