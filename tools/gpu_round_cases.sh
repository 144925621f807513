NO_NCU=1 bash tools/gpu_round.sh r15
timeout 500 python tools/case_bench.py > gpurun_out/r15_cases.jsonl 2> gpurun_out/r15_cases.err; tail -3 gpurun_out/r15_cases.err; cat gpurun_out/r15_cases.jsonl | cut -c1-330
