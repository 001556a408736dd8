#!/bin/bash
N=${1:-8}
mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/r2_bench_${N}gpu.json 2> gpurun_out/r2_bench_${N}gpu.err
echo rc=$?
tail -3 gpurun_out/r2_bench_${N}gpu.err | cut -c1-300
python - <<PY
import json
d=json.load(open('gpurun_out/r2_bench_${N}gpu.json'))
print("value", d["value"], "ms", d["ms_per_step"], "e2e", d["e2e"]["value"], "fp32 e2e", d["e2e_fp32_vectors"]["value"], "clocks", d["clocks"])
s=d["slq_logdet"]; print("slq", s["seconds"], s["layout"], s["launches_per_logdet"], s["roofline"]["frac"], "lanczos", s["lanczos_form"]["seconds"], s["logdet_estimate"])
print(json.dumps(d["extra_workloads"], indent=1)[:2500])
PY
