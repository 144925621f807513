set -x
mkdir -p gpurun_out
python tools/pcie_probe.py > gpurun_out/s7_pcie.json 2>&1; cat gpurun_out/s7_pcie.json
timeout 1500 python bench.py --size 256 --steps 6 --warmup 3 --no-cpu-baseline > gpurun_out/s7_bench256.json 2> gpurun_out/s7_bench256.err; tail -3 gpurun_out/s7_bench256.err
python - <<PY
import json
d=json.loads(open('gpurun_out/s7_bench256.json').read().strip().split('\n')[-1])
print({k:d.get(k) for k in ('value','e2e','e2e_blocks','wall_s')})
PY
