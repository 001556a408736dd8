mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_golden_fixtures.py tests/test_gpu_zgrad.py -m gpu -x -q > gpurun_out/pytest_skinny.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_skinny.log
LIP_SKINNY_ROWTHREAD=0 python bench.py --workload lenet5 --no-cpu --no-e2e --steps 5 --warmup 3 > gpurun_out/bench_lenet5_old.json 2>/dev/null
python bench.py --workload lenet5 --no-cpu --no-e2e --steps 5 --warmup 3 > gpurun_out/bench_lenet5_new.json 2>/dev/null
for f in old new; do python -c "
import json;d=json.load(open('gpurun_out/bench_lenet5_$f.json'));print('$f',d['value'],d['ms_per_step'],d['roofline']['ms_per_call'])"; done
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_lenet5.csv python bench.py --workload lenet5 --no-cpu --no-e2e --steps 1 --warmup 1 > gpurun_out/ncu_lenet.log 2>&1
python tools/summarize_launches.py gpurun_out/launches_lenet5.csv 14 | tee gpurun_out/launches_lenet5_summary.txt
