set -x
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q -k "materialize or golden" > gpurun_out/pytest_small.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_small.log
python bench.py --no-cpu > gpurun_out/bench_lz.json 2> gpurun_out/bench_lz.err; tail -3 gpurun_out/bench_lz.err
