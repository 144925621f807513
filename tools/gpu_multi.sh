# multi-GPU round (run under gpurun --gpus N): decomposition parity on NCCL + weak-scaling bench lines
set -x
TAG=${1:-mX}; N=${2:-2}; SZ=${3:-256}
mkdir -p gpurun_out
python -m pytest tests/test_gpu_multirank.py -m gpu -x -q > gpurun_out/${TAG}_pytest.log 2>&1; tail -3 gpurun_out/${TAG}_pytest.log
python bench.py --steps 5 --warmup 3 --size $SZ --no-cpu-baseline > gpurun_out/${TAG}_bench_n1.json 2> gpurun_out/${TAG}_bench_n1.err
for g in 2 4 8; do
  if [ $g -le $N ]; then
    python -m torch.distributed.run --nnodes=1 --nproc-per-node $g --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $g --steps 5 --warmup 3 --size $SZ --no-cpu-baseline > gpurun_out/${TAG}_bench_n$g.json 2> gpurun_out/${TAG}_bench_n$g.err
  fi
done
for f in gpurun_out/${TAG}_bench_n*.json; do python tools/bench_summary.py $f 2>/dev/null | head -3; done
