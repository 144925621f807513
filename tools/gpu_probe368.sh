mkdir -p gpurun_out
python bench.py --size 320 --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/r12_bench320.json 2> gpurun_out/r12_bench320.err
python tools/bench_summary.py gpurun_out/r12_bench320.json 2>/dev/null | head -6
ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:'FluxTileBody|FluxGradTileBody' -s 4 -c 2 -o gpurun_out/r12_tiles368 python tools/run_step.py --n 368 --steps 1 > gpurun_out/r12_ncu368.log 2>&1
tail -3 gpurun_out/r12_ncu368.log
