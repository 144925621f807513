# round 2, first GPU call: host topology, parity tests, smoke, default bench (both arms)
set -x
TAG=${1:-s1}
mkdir -p gpurun_out
{ nproc; lscpu | grep -E "Model name|Socket|NUMA|Thread|Core"; cat /sys/fs/cgroup/cpuset.cpus.effective /sys/fs/cgroup/cpuset.mems.effective; free -g | head -2;
  nvidia-smi topo -m; for d in /sys/bus/pci/devices/*; do if [ -f $d/numa_node ] && grep -qi "0x10de" $d/vendor 2>/dev/null; then echo $d $(cat $d/numa_node) $(cat $d/class); fi; done; } > gpurun_out/${TAG}_host.txt 2>&1
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/${TAG}_pytest.log 2>&1; tail -5 gpurun_out/${TAG}_pytest.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/${TAG}_smoke.log 2>&1; tail -2 gpurun_out/${TAG}_smoke.log
timeout 600 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/${TAG}_bench_reference.json 2> gpurun_out/${TAG}_bench_reference.err; tail -c 1500 gpurun_out/${TAG}_bench_reference.json
timeout 1200 python bench.py > gpurun_out/${TAG}_bench_default.json 2> gpurun_out/${TAG}_bench_default.err; tail -5 gpurun_out/${TAG}_bench_default.err
python tools/bench_summary.py gpurun_out/${TAG}_bench_default.json | head -30
