# final evidence of a round: parity tests, smoke, bench lines (128^3 fp64/fp32, 256^3), launch list + full ncu of the hot kernels,
# then the default bench (368^3) exactly as the driver runs it
TAG=${1:-rF}
bash tools/gpu_round.sh $TAG
(time timeout 900 python bench.py > gpurun_out/${TAG}_bench_default.json 2> gpurun_out/${TAG}_bench_default.err); tail -4 gpurun_out/${TAG}_bench_default.err
python tools/bench_summary.py gpurun_out/${TAG}_bench_default.json 2>/dev/null | head -7
