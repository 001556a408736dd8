#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_golden_fixtures.py tests/test_gpu_objectives.py -m gpu -q 2>&1 | tail -4
for f in 0 1; do
  LIP_HEAD_FUSE=$f timeout 300 python bench.py --no-slq --no-extra --no-train-step --no-cpu --steps 20 --warmup 5 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); r=d['roofline']
print('LIP_HEAD_FUSE=$f', 'value', round(d['value']), 'e2e', round(d['e2e']['value']), 'ms/call', round(r['ms_per_call'],3), 'gauss', round(r['ms_per_call_gaussian_probes'],3), 'frac', round(r['frac'],3), round(r['frac_gaussian_probes'],3), 'launches', d['gpu_launches'], 'trace', d['hutchinson_trace_estimate'])"
done
