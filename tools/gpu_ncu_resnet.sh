set -x
mkdir -p gpurun_out
python bench.py --workload resnet1m --probes 16 --no-cpu --no-e2e --steps 1 --warmup 1 > gpurun_out/plain_resnet.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv --log-file gpurun_out/launches_resnet1m.csv \
  python bench.py --workload resnet1m --probes 16 --no-cpu --no-e2e --steps 1 --warmup 1 > gpurun_out/ncu_resnet.log 2>&1
python tools/summarize_launches.py gpurun_out/launches_resnet1m.csv 25 > gpurun_out/launches_resnet1m_summary.txt; head -14 gpurun_out/launches_resnet1m_summary.txt
# full sections of the dominant conv kernels (a handful of launches of the second call)
ncu --set full --clock-control none --import-source on -k regex:conv_tc_kernel --launch-skip 120 --launch-count 8 -f -o gpurun_out/r01_conv_tc \
  python bench.py --workload resnet1m --probes 16 --no-cpu --no-e2e --steps 1 --warmup 1 > gpurun_out/ncu_resnet_full.log 2>&1
# zgrad launch list (the f1 kernels)
python tools/zgrad_time.py 512 64 256 409 > /dev/null 2>&1
ls -la gpurun_out | tail
