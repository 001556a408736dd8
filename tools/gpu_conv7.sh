mkdir -p gpurun_out
for m in 1 0; do
echo "== MERGE=$m"
LIP_CONV_TC_MERGE=$m timeout 300 python tools/conv_tc_selftest.py 10 big 2>&1 | grep "imgs=100" | grep "role 0\|role 2" | grep "stride=1"
done > gpurun_out/conv_merge_ab.log 2>&1
cat gpurun_out/conv_merge_ab.log
