# round 2: full parity suite, smoke, default bench (both arms), DRAM-byte capture at the bench size
set -x
TAG=${1:-s5}
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/${TAG}_pytest.log 2>&1; tail -5 gpurun_out/${TAG}_pytest.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/${TAG}_smoke.log 2>&1; tail -2 gpurun_out/${TAG}_smoke.log
timeout 600 python bench.py --impl reference > gpurun_out/${TAG}_bench_reference.json 2> gpurun_out/${TAG}_bench_reference.err; cut -c1-300 gpurun_out/${TAG}_bench_reference.json
timeout 1500 python bench.py > gpurun_out/${TAG}_bench_default.json 2> gpurun_out/${TAG}_bench_default.err; tail -6 gpurun_out/${TAG}_bench_default.err
python tools/bench_summary.py gpurun_out/${TAG}_bench_default.json > gpurun_out/${TAG}_bench_default_summary.txt; head -8 gpurun_out/${TAG}_bench_default_summary.txt
python - <<PY
import json
d=json.loads(open('gpurun_out/${TAG}_bench_default.json').read().strip().split('\n')[-1])
print({k:d.get(k) for k in ('value','e2e','parity_maxerr','wall_s')}); print(d.get('fp32',{}).get('value'), d['roofline']['stage_model'])
PY
if [ -n "$NCU368" ]; then
ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none --kernel-name-base demangled -k regex:'FluxTileBody|FluxGradTileBody|GradAdjTileBody|GradCellBody' -s 20 -c 17 --csv --log-file gpurun_out/${TAG}_dram368.csv python tools/run_step.py --n 368 --steps 1 > gpurun_out/${TAG}_ncu368.log 2>&1
tail -3 gpurun_out/${TAG}_ncu368.log
fi
du -sh gpurun_out
