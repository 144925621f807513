# kernel experiments: second builds of the library under exp/ selected with ADFVM_B200_LIB
set -x
mkdir -p gpurun_out
for v in $VARIANTS; do
  export ADFVM_B200_LIB=$PWD/exp/lib_$v.so
  if [ -z "$NO_TESTS" ]; then timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "large_parity_against_reference or bitwise or golden_adjoint" > gpurun_out/x_${v}_pytest.log 2>&1; tail -2 gpurun_out/x_${v}_pytest.log; fi
  for n in $SIZES; do
    timeout 600 python bench.py --steps 6 --warmup 3 --size $n --no-cpu-baseline > gpurun_out/x_${v}_bench$n.json 2> gpurun_out/x_${v}_bench$n.err
    python tools/bench_summary.py gpurun_out/x_${v}_bench$n.json 2>/dev/null | head -7
  done
done
unset ADFVM_B200_LIB
