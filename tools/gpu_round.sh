set -x
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/r7_pytest.log 2>&1; tail -3 gpurun_out/r7_pytest.log
python __graft_entry__.py smoke > gpurun_out/r7_smoke.log 2>&1; tail -2 gpurun_out/r7_smoke.log
python bench.py --steps 10 --warmup 3 > gpurun_out/r7_bench128_f64.json 2> gpurun_out/r7_bench128_f64.err
python bench.py --steps 10 --warmup 3 --dtype f32 --no-cpu-baseline > gpurun_out/r7_bench128_f32.json 2> gpurun_out/r7_bench128_f32.err
python bench.py --steps 5 --warmup 3 --size 256 --no-cpu-baseline > gpurun_out/r7_bench256_f64.json 2> gpurun_out/r7_bench256_f64.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r7_launches_f64.csv python tools/run_step.py --n 128 --steps 2 > gpurun_out/r7_launch.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_tile -s 12 -c 4 -o gpurun_out/r7_tiles_f64 python tools/run_step.py --n 128 --steps 2 > gpurun_out/r7_ncu1.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_tile -s 12 -c 4 -o gpurun_out/r7_tiles_f32 python tools/run_step.py --n 128 --steps 2 --dtype f32 > gpurun_out/r7_ncu2.log 2>&1
cat gpurun_out/r7_bench128_f64.json | cut -c1-1500
