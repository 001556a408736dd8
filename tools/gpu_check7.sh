set -x
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -3 gpurun_out/pytest_gpu.log
python bench.py > gpurun_out/bench_full.json 2> gpurun_out/bench_full.err; cat gpurun_out/bench_full.json; tail -3 gpurun_out/bench_full.err
python tools/slq_time.py 4 40 > gpurun_out/plain_slq.log 2>&1 && \
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,dram__throughput.avg.pct_of_peak_sustained_elapsed --clock-control none -k regex:"reorth|basis_axpy|dot_partial" --launch-skip 200 -c 40 --csv --log-file gpurun_out/vecstage_slq.csv \
  python tools/slq_time.py 4 40 > gpurun_out/ncu_slq.log 2>&1
