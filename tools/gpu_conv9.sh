mkdir -p gpurun_out
for v in "LIP_CONV_TC_MERGE=1" "LIP_CONV_TC_MERGE=1 LIP_SELFTEST_NOADD=1"; do
echo "== $v"
env $v timeout 300 python tools/conv_tc_selftest.py 10 big 2>&1 | grep "imgs=100" | grep "role 2" | cut -c1-140
done > gpurun_out/conv_merge_ab3.log 2>&1
cat gpurun_out/conv_merge_ab3.log
python bench.py --workload resnet1m --probes 16 --no-cpu --no-e2e --steps 3 --warmup 1 > gpurun_out/bench_resnet1m.json 2> gpurun_out/bench_resnet1m.err; cut -c1-200 gpurun_out/bench_resnet1m.json; tail -5 gpurun_out/bench_resnet1m.err
