mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_golden_fixtures.py tests/test_gpu_zgrad.py tests/test_gpu_objectives.py -m gpu -x -q > gpurun_out/pytest_skinny.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_skinny.log
for cfg in "0 0" "0 1" "1 0" "1 1"; do set -- $cfg
LIP_SKINNY_ROWTHREAD=$1 LIP_FOLD_N=$2 python bench.py --workload lenet5 --no-cpu --no-e2e --steps 5 --warmup 3 > gpurun_out/bench_lenet5_$1$2.json 2>/dev/null
python -c "
import json;d=json.load(open('gpurun_out/bench_lenet5_$1$2.json'));print('rowthread=$1 fold=$2',d['value'],d['ms_per_step'],d['roofline']['ms_per_call'])"; done
LIP_SKINNY_ROWTHREAD=0 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_lenet5.csv python bench.py --workload lenet5 --no-cpu --no-e2e --steps 1 --warmup 1 > gpurun_out/ncu_lenet.log 2>&1
python tools/summarize_launches.py gpurun_out/launches_lenet5.csv 14 | tee gpurun_out/launches_lenet5_summary.txt
