#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -k "lenet or cnn or conv" 2>&1 | tail -8
for f in 1 0; do
LIP_CNN_FUSE=$f timeout 300 python bench.py --workload lenet5 --no-cpu --no-e2e --no-extra --steps 5 --warmup 3 2>gpurun_out/lenet_fuse$f.err | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('fuse=$f value', round(d['value']), 'ms', round(d['ms_per_step'],3), 'launches', d['gpu_launches'], 'trace', d.get('hutchinson_trace_estimate'))"
done
