set -x
mkdir -p gpurun_out
python bench.py --no-cpu --no-slq --no-e2e --no-train-step --steps 2 --warmup 1 > gpurun_out/plain_final.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_bench_final.csv python bench.py --no-cpu --no-slq --no-e2e --no-train-step --steps 2 --warmup 1 > gpurun_out/ncu_lf.log 2>&1
python tools/summarize_launches.py gpurun_out/launches_bench_final.csv 16 | tee gpurun_out/launches_bench_final_summary.txt
