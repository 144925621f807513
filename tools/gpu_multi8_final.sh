# 8-GPU evidence: NCCL decomposition parity on 8 ranks (incl. adjoint viscosity), then the driver's 8-GPU bench command at the default size
set -x
mkdir -p gpurun_out
free -g | head -2 > gpurun_out/m8f_host.txt; nproc >> gpurun_out/m8f_host.txt; nvidia-smi topo -m >> gpurun_out/m8f_host.txt 2>&1
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 tools/multigpu_check.py > gpurun_out/m8f_check.log 2>&1; grep maxerr gpurun_out/m8f_check.log
(time timeout 800 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 8 --steps 3 --warmup 3 > gpurun_out/m8f_bench_n8.json 2> gpurun_out/m8f_bench_n8.err)
tail -4 gpurun_out/m8f_bench_n8.err
python tools/bench_summary.py gpurun_out/m8f_bench_n8.json 2>/dev/null | head -6
python - <<PY
import json
d=json.loads([l for l in open('gpurun_out/m8f_bench_n8.json') if l.startswith('{')][-1])
print({k:d.get(k) for k in ('value','e2e','e2e_blocks','parity','strong','wall_s')})
PY
