set -x
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q -k "resnet or lenet" > gpurun_out/pytest_cnn.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_cnn.log
tail -15 gpurun_out/pytest_cnn.log
python bench.py --workload resnet1m --probes 16 --no-cpu --no-e2e --steps 2 --warmup 1 > gpurun_out/bench_resnet1m.json 2> gpurun_out/bench_resnet1m.err; cut -c1-250 gpurun_out/bench_resnet1m.json; tail -3 gpurun_out/bench_resnet1m.err
