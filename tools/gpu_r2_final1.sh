#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -3
timeout 900 python bench.py > gpurun_out/r2_final_1gpu.json 2> gpurun_out/r2_final_1gpu.err; echo rc=$?
tail -3 gpurun_out/r2_final_1gpu.err | cut -c1-300
timeout 600 python bench.py --impl reference > gpurun_out/r2_final_ref.json 2> gpurun_out/r2_final_ref.err; echo rc=$?
cat gpurun_out/r2_final_ref.json | cut -c1-600
python - <<PY
import json
d=json.load(open('gpurun_out/r2_final_1gpu.json'))
print("value", d["value"], "ms", d["ms_per_step"], "e2e", d["e2e"]["value"], "fp32 e2e", d["e2e_fp32_vectors"]["value"], "clocks", d["clocks"])
print("roofline", d["roofline"]); print("cpu", d["cpu_baseline"])
s=d["slq_logdet"]; print("slq", s["seconds"], s["layout"], s["launches_per_logdet"], s["roofline"]["frac"], "lanczos", s["lanczos_form"]["seconds"], s["logdet_estimate"])
print(json.dumps(d["extra_workloads"], indent=1)[:2500])
PY
