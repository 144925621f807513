NO_NCU=1 bash tools/gpu_round.sh r12
(time python bench.py > gpurun_out/r12_bench_default.json 2> gpurun_out/r12_bench_default.err); tail -4 gpurun_out/r12_bench_default.err
python tools/bench_summary.py gpurun_out/r12_bench_default.json 2>/dev/null | head -7
