# the driver's 8-GPU bench command at the default size (368^3 per GPU when the host memory allows it)
mkdir -p gpurun_out
free -g | head -2; nproc
(time timeout 800 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 8 --steps 3 --warmup 3 > gpurun_out/m8d_bench_n8.json 2> gpurun_out/m8d_bench_n8.err)
tail -4 gpurun_out/m8d_bench_n8.err
python tools/bench_summary.py gpurun_out/m8d_bench_n8.json 2>/dev/null | head -4
python -c "
import json; d=json.loads([l for l in open('gpurun_out/m8d_bench_n8.json') if l.startswith('{')][-1]); print(d['config'])"
