set -x
mkdir -p gpurun_out
timeout 300 python tools/conv_tc_selftest.py 5 > gpurun_out/conv_selftest.log 2>&1
echo "selftest rc=$?"; grep -n "imgs=100\|ALL OK\|FAIL\|rc=-" gpurun_out/conv_selftest.log
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -s -k "resnet or conv or lenet" > gpurun_out/pytest_resnet.log 2>&1
echo "pytest rc=$?"; grep -n "rel err\|passed\|failed\|Error\|assert" gpurun_out/pytest_resnet.log | tail -15
python bench.py --workload resnet1m --probes 16 --no-cpu --no-e2e --steps 3 --warmup 1 > gpurun_out/bench_resnet1m.json 2> gpurun_out/bench_resnet1m.err; cat gpurun_out/bench_resnet1m.json; tail -5 gpurun_out/bench_resnet1m.err
LIP_CONV_TC_FOLD=0 python bench.py --workload resnet1m --probes 16 --no-cpu --no-e2e --steps 3 --warmup 1 > gpurun_out/bench_resnet1m_nofold.json 2> gpurun_out/bench_resnet1m_nofold.err; cat gpurun_out/bench_resnet1m_nofold.json
ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file gpurun_out/launches_resnet1m.csv python bench.py --workload resnet1m --probes 16 --no-cpu --no-e2e --steps 1 --warmup 1 > gpurun_out/ncu_resnet.log 2>&1
