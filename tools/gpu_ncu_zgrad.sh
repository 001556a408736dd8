set -x
mkdir -p gpurun_out
python tools/zgrad_time.py 512 64 256 409 > /dev/null 2>&1
# full sections of the tcgen05 GEMMs of one lip_zgrad call (the third timed call) + the elementwise reverse kernel
ncu --set full --clock-control none --import-source on -k regex:"gemm_tc|reverse_act_tc" --launch-skip 60 --launch-count 14 -f -o gpurun_out/r01_zgrad_tc \
  python tools/zgrad_time.py 512 64 256 409 > gpurun_out/ncu_zgrad_full.log 2>&1
ls -la gpurun_out/*.ncu-rep
