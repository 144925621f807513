# last check of a round: parity + smoke + the default bench line exactly as the driver runs it
TAG=${1:-rL}
mkdir -p gpurun_out
timeout 400 python -m pytest tests -m gpu -x -q > gpurun_out/${TAG}_pytest.log 2>&1; tail -2 gpurun_out/${TAG}_pytest.log
timeout 200 python __graft_entry__.py smoke > gpurun_out/${TAG}_smoke.log 2>&1; tail -1 gpurun_out/${TAG}_smoke.log
(time timeout 900 python bench.py > gpurun_out/${TAG}_bench_default.json 2> gpurun_out/${TAG}_bench_default.err); tail -3 gpurun_out/${TAG}_bench_default.err
python tools/bench_summary.py gpurun_out/${TAG}_bench_default.json 2>/dev/null | head -7
