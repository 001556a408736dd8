set -x
mkdir -p gpurun_out
python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/smoke.log; tail -2 gpurun_out/smoke.log
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -3 gpurun_out/pytest_gpu.log
python bench.py > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err; echo "bench rc=$?"
python tools/zgrad_time.py 512 64 256 409 > gpurun_out/zgrad_time.txt 2>&1
python tools/zgrad_time.py 50 64 256 40 >> gpurun_out/zgrad_time.txt 2>&1
cat gpurun_out/zgrad_time.txt
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"gemm|zgrad|reverse_act|rho|act_second|batch_sum|tf32_split|bias|factor" --launch-skip 0 -c 200 --csv --log-file gpurun_out/launches_zgrad.csv python tools/zgrad_time.py 512 64 256 409 > gpurun_out/ncu_zgrad.log 2>&1
python tools/summarize_launches.py gpurun_out/launches_zgrad.csv 16 > gpurun_out/launches_zgrad_summary.txt; cat gpurun_out/launches_zgrad_summary.txt
python - <<'PY'
import json
d=json.load(open("gpurun_out/bench_default.json"))
print(d["value"], d["roofline"]["frac"], d["e2e"]["value"], d["slq_logdet"]["seconds"], d["train_step"]["zgrad_ms"], d["train_step"]["optimize_step_seconds"], d["clocks"])
PY
