mkdir -p gpurun_out
( cd tmp_v1 && timeout 300 python tools/conv_tc_selftest.py 10 2>&1 | grep "imgs=100" | grep "role 2" ) > gpurun_out/conv_v1_again.log 2>&1
echo "== current" >> gpurun_out/conv_v1_again.log
timeout 300 python tools/conv_tc_selftest.py 10 big 2>&1 | grep "role 2" | grep "stride=1" >> gpurun_out/conv_v1_again.log
cat gpurun_out/conv_v1_again.log
