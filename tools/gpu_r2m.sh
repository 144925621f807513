# round 2 multi-GPU check: NCCL decomposition parity (48^3 per rank, early tiles > 0), multi-rank pytest, both bench arms under torchrun
set -x
N=${1:-2}; TAG=${2:-m1}
mkdir -p gpurun_out
free -g | head -2 > gpurun_out/${TAG}_host.txt; nvidia-smi topo -m >> gpurun_out/${TAG}_host.txt 2>&1; nproc >> gpurun_out/${TAG}_host.txt
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 tools/multigpu_check.py > gpurun_out/${TAG}_check.log 2>&1; grep maxerr gpurun_out/${TAG}_check.log
timeout 900 python -m pytest tests/test_gpu_multirank.py -m gpu -x -q > gpurun_out/${TAG}_pytest.log 2>&1; tail -3 gpurun_out/${TAG}_pytest.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus $N --steps 3 --warmup 1 > gpurun_out/${TAG}_bench_reference.json 2> gpurun_out/${TAG}_bench_reference.err; cut -c1-400 gpurun_out/${TAG}_bench_reference.json
timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus $N ${BENCH_ARGS} > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; grep "bench " gpurun_out/${TAG}_bench.err | tail -12
python tools/bench_summary.py gpurun_out/${TAG}_bench.json > gpurun_out/${TAG}_bench_summary.txt; head -8 gpurun_out/${TAG}_bench_summary.txt
python - <<PY
import json
d=json.loads(open('gpurun_out/${TAG}_bench.json').read().strip().split('\n')[-1])
print({k:d.get(k) for k in ('value','e2e','parity','strong','wall_s')})
PY
