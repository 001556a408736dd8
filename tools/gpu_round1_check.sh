set -x
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
LIP_TC_2CTA=1 python -m pytest tests -m gpu -x -q -k "tensor_core or headline" > gpurun_out/pytest_gpu_2cta.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu_2cta.log
python bench.py > gpurun_out/bench_1cta.json 2> gpurun_out/bench_1cta.err
LIP_TC_2CTA=1 python bench.py --no-cpu > gpurun_out/bench_2cta.json 2> gpurun_out/bench_2cta.err
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err
nvidia-smi > gpurun_out/nvidia_smi.txt; nproc >> gpurun_out/nvidia_smi.txt
tail -3 gpurun_out/pytest_gpu.log gpurun_out/pytest_gpu_2cta.log
cat gpurun_out/bench_1cta.json gpurun_out/bench_2cta.json gpurun_out/bench_ref.json
