mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/s13_pytest.log 2>&1; tail -2 gpurun_out/s13_pytest.log
timeout 600 python bench.py --steps 6 --warmup 3 --size 256 --no-cpu-baseline > gpurun_out/s13_bench256.json 2> gpurun_out/s13_bench256.err
python tools/bench_summary.py gpurun_out/s13_bench256.json | head -7
timeout 600 python bench.py --steps 10 --warmup 3 --size 128 --no-cpu-baseline > gpurun_out/s13_bench128.json 2> gpurun_out/s13_bench128.err
python tools/bench_summary.py gpurun_out/s13_bench128.json | head -7
