#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -k "lenet" 2>&1 | tail -3
for v in "0 16" "1 16" "0 8" "1 8" "1 5"; do
set -- $v
LIP_CNN_VJP_VARIANT=$1 LIP_CNN_VJP_GMUL=$2 timeout 300 python bench.py --workload lenet5 --no-cpu --no-e2e --no-extra --steps 5 --warmup 3 2>gpurun_out/lenet_v.err | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('variant $1 gmul $2 value', round(d['value']), 'ms', round(d['ms_per_step'],3))"
done
