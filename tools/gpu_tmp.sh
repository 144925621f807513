NO_TESTS=1 VARIANTS="co0 co1" SIZES="256" bash tools/gpu_exp.sh
