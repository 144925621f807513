VARIANTS="pf0 pf1" SIZES="256" bash tools/gpu_exp.sh
