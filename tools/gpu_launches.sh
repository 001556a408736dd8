set -x
mkdir -p gpurun_out
python bench.py --no-cpu --no-slq --no-e2e --steps 2 --warmup 1 > gpurun_out/plain_1cta.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_1cta.csv python bench.py --no-cpu --no-slq --no-e2e --steps 2 --warmup 1 > gpurun_out/ncu_l1.log 2>&1
LIP_TC_2CTA=1 python bench.py --no-cpu --no-slq --no-e2e --steps 2 --warmup 1 > gpurun_out/plain_2cta.log 2>&1 && \
LIP_TC_2CTA=1 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_2cta.csv python bench.py --no-cpu --no-slq --no-e2e --steps 2 --warmup 1 > gpurun_out/ncu_l2.log 2>&1
