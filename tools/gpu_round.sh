# one GPU round: parity tests, smoke, bench lines, launch list, full ncu capture of the hot kernels (run under gpurun)
set -x
TAG=${1:-rX}
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/${TAG}_pytest.log 2>&1; tail -3 gpurun_out/${TAG}_pytest.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/${TAG}_smoke.log 2>&1; tail -2 gpurun_out/${TAG}_smoke.log
timeout 600 python bench.py --steps 10 --warmup 3 --size 128 > gpurun_out/${TAG}_bench128_f64.json 2> gpurun_out/${TAG}_bench128_f64.err
timeout 600 python bench.py --steps 10 --warmup 3 --size 128 --dtype f32 --no-cpu-baseline > gpurun_out/${TAG}_bench128_f32.json 2> gpurun_out/${TAG}_bench128_f32.err
timeout 600 python bench.py --steps 5 --warmup 3 --size 256 --no-cpu-baseline > gpurun_out/${TAG}_bench256_f64.json 2> gpurun_out/${TAG}_bench256_f64.err
if [ -z "$NO_NCU" ]; then
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${TAG}_launches_f64.csv python tools/run_step.py --n 128 --steps 2 > gpurun_out/${TAG}_launch.log 2>&1
ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:'FluxTileBody|FluxGradTileBody|GradAdjUpdateBody' -s 6 -c 6 -o gpurun_out/${TAG}_tiles_f64 python tools/run_step.py --n 128 --steps 2 > gpurun_out/${TAG}_ncu1.log 2>&1
ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:'FluxTileBody|FluxGradTileBody|GradAdjUpdateBody' -s 6 -c 6 -o gpurun_out/${TAG}_tiles_f32 python tools/run_step.py --n 128 --steps 2 --dtype f32 > gpurun_out/${TAG}_ncu2.log 2>&1
fi
for f in gpurun_out/${TAG}_bench*.json; do python tools/bench_summary.py $f | head -8; done
