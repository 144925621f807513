# round 2 kernel iteration: parity tests, quick benches, ncu captures (launch list, full set at 128^3, DRAM bytes at 368^3)
set -x
TAG=${1:-s2}
mkdir -p gpurun_out
if [ -z "$NO_TESTS" ]; then
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/${TAG}_pytest.log 2>&1; tail -5 gpurun_out/${TAG}_pytest.log
fi
timeout 600 python bench.py --steps 10 --warmup 3 --size 128 --no-cpu-baseline --no-parity > gpurun_out/${TAG}_bench128.json 2> gpurun_out/${TAG}_bench128.err
timeout 600 python bench.py --steps 5 --warmup 3 --size 256 --no-cpu-baseline --no-parity --no-fp32 > gpurun_out/${TAG}_bench256.json 2> gpurun_out/${TAG}_bench256.err
for f in gpurun_out/${TAG}_bench128.json gpurun_out/${TAG}_bench256.json; do python tools/bench_summary.py $f > gpurun_out/${TAG}_summary_$(basename $f .json).txt; head -9 gpurun_out/${TAG}_summary_$(basename $f .json).txt; done
if [ -z "$NO_NCU" ]; then
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${TAG}_launches_f64.csv python tools/run_step.py --n 128 --steps 2 > gpurun_out/${TAG}_launch.log 2>&1
# one resident primal + adjoint step = 41 launches; skip the priming calls and the first resident step, capture the next one
ncu --set full --clock-control none --kernel-name-base demangled -k regex:'k_tile|k_run|k_reduce' -s 84 -c 41 -o /tmp/${TAG}_all_f64 python tools/run_step.py --n 128 --steps 2 > gpurun_out/${TAG}_ncu1.log 2>&1
python tools/ncu_summary.py /tmp/${TAG}_all_f64.ncu-rep > gpurun_out/${TAG}_all_f64.summary.txt 2>&1
ls -la /tmp/${TAG}_all_f64.ncu-rep
ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:'FluxGradTileBody|GradAdjTileBody' -s 6 -c 2 -o gpurun_out/${TAG}_tiles_f64 python tools/run_step.py --n 128 --steps 2 > gpurun_out/${TAG}_ncu2.log 2>&1
fi
if [ -n "$NCU368" ]; then
ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none --kernel-name-base demangled -k regex:'FluxTileBody|FluxGradTileBody|GradAdjTileBody|GradCellBody' -s 20 -c 17 --csv --log-file gpurun_out/${TAG}_dram368.csv python tools/run_step.py --n 368 --steps 1 > gpurun_out/${TAG}_ncu368.log 2>&1
fi
du -sh gpurun_out; ls -la gpurun_out | tail -14
