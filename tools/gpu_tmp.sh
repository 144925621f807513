mkdir -p gpurun_out
ADFVM_TILE_TIMING=1 python tools/setup_time.py --n 256 > gpurun_out/s8_setup256.log 2>&1; cat gpurun_out/s8_setup256.log; nproc
