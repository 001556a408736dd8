mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_zgrad.py tests/test_gpu_objectives.py tests/test_golden_fixtures.py -m gpu -q -s 2>&1 | grep -v "Warning\|warnings.warn" > gpurun_out/pytest_zgrad_tc.log; grep -n "rel err\|passed\|failed\|Error\|error" gpurun_out/pytest_zgrad_tc.log | head -30
timeout 600 python tools/zgrad_time.py 512 64 256 409 2>&1 | grep zgrad | tail -3
