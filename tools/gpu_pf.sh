mkdir -p gpurun_out
LIP_TC_WIDE=1 timeout 300 python tools/tc_selftest.py > gpurun_out/selftest_wide.log 2>&1; tail -1 gpurun_out/selftest_wide.log
: > gpurun_out/pf_sweep.log
for PF in 0 2 4 8 16; do
  LIP_TC_PF=$PF python bench.py --no-cpu --no-slq --no-e2e --steps 20 --warmup 3 2>/dev/null | python -c "
import json,sys,os
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); r=d['roofline']
print('LIP_TC_PF=%s  %9.0f products/s  %.3f ms/step  lip_ggn_vp %.3f ms  gaussian %.3f ms  clocks %s' % (os.environ.get('LIP_TC_PF'), d['value'], d['ms_per_step'], r['ms_per_call'], r['ms_per_call_gaussian_probes'], d['clocks']['sm_mhz']))
" >> gpurun_out/pf_sweep.log
done
cat gpurun_out/pf_sweep.log
