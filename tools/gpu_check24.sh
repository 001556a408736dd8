mkdir -p gpurun_out
python bench.py > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err; echo "bench rc=$?"; tail -3 gpurun_out/bench_default.err
python - <<'PY'
import json
d=json.load(open("gpurun_out/bench_default.json"))
print(d["value"], d["roofline"]["frac"], d["e2e"]["value"], d["slq_logdet"]["seconds"], d["train_step"])
PY
timeout 900 python bench.py --workload resnet1m --points 4096 --probes 4 --no-cpu --no-e2e --steps 3 --warmup 3 > gpurun_out/bench_resnet1m_m4096.json 2> gpurun_out/bench_resnet1m_m4096.err; echo "rc=$?"; cut -c1-300 gpurun_out/bench_resnet1m_m4096.json; tail -3 gpurun_out/bench_resnet1m_m4096.err
