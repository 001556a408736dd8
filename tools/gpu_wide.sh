set -x
mkdir -p gpurun_out
LIP_TC_WIDE=1 timeout 300 python tools/tc_selftest.py > gpurun_out/selftest_wide.log 2>&1; echo "selftest rc=$?"; tail -30 gpurun_out/selftest_wide.log
LIP_BENCH_RANDOM=1 timeout 300 python tools/tc_microbench.py 20 2>&1 | grep "dbg=0\|dbg=1\|ERROR" > gpurun_out/microbench_wide.log; cat gpurun_out/microbench_wide.log
