set -x
mkdir -p gpurun_out
python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/smoke.log; tail -2 gpurun_out/smoke.log
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -3 gpurun_out/pytest_gpu.log
python bench.py > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err; echo "bench rc=$?"
python bench.py --workload lenet5 --no-cpu --steps 5 --warmup 3 > gpurun_out/bench_lenet5.json 2> gpurun_out/bench_lenet5.err
python bench.py --workload resnet1m --probes 16 --no-cpu --no-e2e --steps 5 --warmup 3 > gpurun_out/bench_resnet1m.json 2> gpurun_out/bench_resnet1m.err
python tools/zgrad_time.py 50 64 256 40 > gpurun_out/zgrad_time_m50.txt 2>&1; python tools/zgrad_time.py 512 64 256 409 > gpurun_out/zgrad_time_m512.txt 2>&1
python tools/lenet_step_time.py > gpurun_out/lenet_step_time.txt 2>&1
python - <<'PY'
import json
for f in ("bench_default","bench_lenet5","bench_resnet1m"):
    d=json.load(open(f"gpurun_out/{f}.json"))
    print(f, d["value"], d["roofline"]["frac"], d.get("e2e") and d["e2e"]["value"], d.get("slq_logdet") and d["slq_logdet"]["seconds"], d.get("train_step"), d["clocks"]["sm_mhz"], d["clocks"]["reasons"])
PY
