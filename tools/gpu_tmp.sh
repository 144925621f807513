VARIANTS="ep0 ep1" SIZES="256" bash tools/gpu_exp.sh
