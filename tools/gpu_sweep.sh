mkdir -p gpurun_out
: > gpurun_out/probe_sweep.log
for B in 64 128 256 512 1024; do
  python bench.py --probes $B --no-cpu --no-slq --no-e2e --steps 10 --warmup 3 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
r=d['roofline']
print('probes/GPU=%4d  %9.0f products/s  %.3f ms/step  lip_ggn_vp %.3f ms (frac %.3f)  gaussian probes %.3f ms (frac %.3f)  clocks %s' % (d['config']['probes_per_gpu'], d['value'], d['ms_per_step'], r['ms_per_call'], r['frac'], r['ms_per_call_gaussian_probes'], r['frac_gaussian_probes'], d['clocks']['sm_mhz']))
" >> gpurun_out/probe_sweep.log
done
cat gpurun_out/probe_sweep.log
