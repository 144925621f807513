mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "visc" > gpurun_out/s12_pytest.log 2>&1; tail -2 gpurun_out/s12_pytest.log
timeout 300 python tools/visc_bench.py --n 256 > gpurun_out/s12_visc256.json 2> gpurun_out/s12_visc256.err; cat gpurun_out/s12_visc256.json; tail -3 gpurun_out/s12_visc256.err
timeout 300 python tools/visc_bench.py --n 256 --dtype f32 --type turkel > gpurun_out/s12_visc256_f32.json 2> gpurun_out/s12_visc256_f32.err; cat gpurun_out/s12_visc256_f32.json
