set -x
mkdir -p gpurun_out
timeout 300 python tools/conv_tc_selftest.py 5 > gpurun_out/conv_selftest.log 2>&1
echo "selftest rc=$?"
tail -40 gpurun_out/conv_selftest.log
