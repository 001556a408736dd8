#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -15 > gpurun_out/r2_tests_all.log
tail -6 gpurun_out/r2_tests_all.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -5
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/r2_bench_1gpu.json 2> gpurun_out/r2_bench_1gpu.err
echo bench rc=$?; tail -3 gpurun_out/r2_bench_1gpu.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/r2_bench_1gpu.json'))
for k in ("value","ms_per_step","config","path","e2e","e2e_fp32_vectors","gpu_launches","cpu_baseline","clocks"):
    print(k, d.get(k))
r=d["roofline"]; print("roofline", {k:r[k] for k in ("achieved","peak","frac","frac_gaussian_probes","ms_per_call","ms_per_call_gaussian_probes")})
s=d["slq_logdet"]; print("slq", {k:s[k] for k in s if k not in ("roofline","lanczos_form")}, s["roofline"]["frac"], s["lanczos_form"]["seconds"])
print("train", d.get("train_step"))
print("extras", json.dumps(d.get("extra_workloads"), indent=1))
PY
timeout 600 python bench.py --impl reference --steps 5 --warmup 2 2>/dev/null | tail -1
