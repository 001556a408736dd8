set -x
mkdir -p gpurun_out
# (a) dominant GEMM kernels of the headline step, full sections (second timed step)
python bench.py --no-cpu --no-slq --no-e2e --steps 2 --warmup 1 > gpurun_out/plain_final.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:gemm_tc --launch-skip 11 --launch-count 11 -f -o gpurun_out/r01_gemm_tc_final \
  python bench.py --no-cpu --no-slq --no-e2e --steps 2 --warmup 1 > gpurun_out/ncu_final.log 2>&1
# (b) DRAM traffic of every launch of one step (for roofline.traffic) + the vector-stage kernels of a short SLQ run
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 200 --csv --log-file gpurun_out/traffic_step.csv \
  python bench.py --no-cpu --no-slq --no-e2e --steps 1 --warmup 1 > gpurun_out/ncu_traffic.log 2>&1
python tools/slq_time.py 4 40 > gpurun_out/plain_slq.log 2>&1 && \
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,dram__throughput.avg.pct_of_peak_sustained_elapsed --clock-control none -k regex:"reorth|basis_axpy|dot_partial|axpby|scale_kernel" --launch-skip 600 -c 60 --csv --log-file gpurun_out/vecstage_slq.csv \
  python tools/slq_time.py 4 40 > gpurun_out/ncu_slq.log 2>&1
ls -la gpurun_out | tail -12
