set -x
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -3 gpurun_out/pytest_gpu.log
python tools/tc_microbench.py 20 > gpurun_out/microbench.log 2>&1
cat gpurun_out/microbench.log
python bench.py --no-cpu --no-slq --no-e2e --steps 2 --warmup 1 > gpurun_out/plain_r01.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:gemm_tc --launch-skip 11 --launch-count 6 -f -o gpurun_out/r01_gemm_tc \
  python bench.py --no-cpu --no-slq --no-e2e --steps 2 --warmup 1 > gpurun_out/ncu_full.log 2>&1
ls -la gpurun_out
