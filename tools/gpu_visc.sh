set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/v1_pytest.log 2>&1; tail -5 gpurun_out/v1_pytest.log
timeout 300 python tools/visc_bench.py --n 128 > gpurun_out/v1_visc128.json 2> gpurun_out/v1_visc128.err; cat gpurun_out/v1_visc128.json; tail -3 gpurun_out/v1_visc128.err
timeout 300 python tools/visc_bench.py --n 256 > gpurun_out/v1_visc256.json 2> gpurun_out/v1_visc256.err; cat gpurun_out/v1_visc256.json; tail -3 gpurun_out/v1_visc256.err
timeout 300 python tools/visc_bench.py --n 256 --dtype f32 --type turkel > gpurun_out/v1_visc256_f32.json 2> gpurun_out/v1_visc256_f32.err; cat gpurun_out/v1_visc256_f32.json
