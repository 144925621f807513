# 8-GPU validation (run under gpurun --gpus 8): decomposition parity on NCCL for 2/4/8 ranks + one weak-scaling bench line at N=8
set -x
TAG=${1:-m8}; SZ=${2:-256}
mkdir -p gpurun_out
python -m pytest tests/test_gpu_multirank.py -m gpu -x -q > gpurun_out/${TAG}_pytest.log 2>&1; tail -3 gpurun_out/${TAG}_pytest.log
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 8 --steps 5 --warmup 3 --size $SZ --no-cpu-baseline > gpurun_out/${TAG}_bench_n8.json 2> gpurun_out/${TAG}_bench_n8.err
tail -5 gpurun_out/${TAG}_bench_n8.err
python tools/bench_summary.py gpurun_out/${TAG}_bench_n8.json 2>/dev/null | head -8
