VARIANTS="acc0 acc1" SIZES="128 256" bash tools/gpu_exp.sh
timeout 300 python tools/visc_bench.py --n 256 > gpurun_out/v1_visc256.json 2> gpurun_out/v1_visc256.err; cat gpurun_out/v1_visc256.json; tail -3 gpurun_out/v1_visc256.err
timeout 300 python tools/visc_bench.py --n 256 --dtype f32 --type turkel > gpurun_out/v1_visc256_f32.json 2> gpurun_out/v1_visc256_f32.err; cat gpurun_out/v1_visc256_f32.json
