#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -6
timeout 300 python bench.py --no-slq --no-extra --no-train-step --no-cpu --steps 20 --warmup 5 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); r=d['roofline']
print('value', round(d['value']), 'e2e', round(d['e2e']['value']), 'fp32 e2e', round(d['e2e_fp32_vectors']['value']), 'ms/call', round(r['ms_per_call'],3), 'gauss', round(r['ms_per_call_gaussian_probes'],3), 'frac', round(r['frac'],3), round(r['frac_gaussian_probes'],3), 'peak', round(r['peak'],1), 'launches', d['gpu_launches'], 'trace', d['hutchinson_trace_estimate'])"
