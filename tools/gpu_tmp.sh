mkdir -p gpurun_out
ADFVM_TILE_TIMING=1 python tools/setup_time.py --n 256 > gpurun_out/s9_setup256.log 2>&1; cat gpurun_out/s9_setup256.log
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "large or golden or bitwise" > gpurun_out/s9_pytest.log 2>&1; tail -2 gpurun_out/s9_pytest.log
