# final evidence of the round: parity suite, smoke, both bench arms at the default size, ncu (launch list, every kernel of one step with
# --set full at 128^3 incl. source for the reverse kernels, DRAM bytes at the bench size, the adjoint-viscosity kernels)
set -x
TAG=${1:-f1}
mkdir -p gpurun_out
NCU368=1 bash tools/gpu_r2c.sh ${TAG}
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${TAG}_launches_f64.csv python tools/run_step.py --n 128 --steps 2 > gpurun_out/${TAG}_launch.log 2>&1
ncu --set full --clock-control none --kernel-name-base demangled -k regex:'k_tile|k_run|k_reduce' -s 84 -c 41 -o /tmp/${TAG}_all_f64 python tools/run_step.py --n 128 --steps 2 > gpurun_out/${TAG}_ncu1.log 2>&1
python tools/ncu_summary.py /tmp/${TAG}_all_f64.ncu-rep > gpurun_out/${TAG}_all_kernels_f64.summary.txt 2>&1
ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:'FluxGradTileBody|GradAdjTileBody' -s 6 -c 2 -o gpurun_out/${TAG}_tiles_f64 python tools/run_step.py --n 128 --steps 2 > gpurun_out/${TAG}_ncu2.log 2>&1
python tools/ncu_summary.py gpurun_out/${TAG}_tiles_f64.ncu-rep > gpurun_out/${TAG}_tiles_f64.summary.txt 2>&1
ncu --set full --clock-control none --kernel-name-base demangled -k regex:'Visc' -s 30 -c 16 -o /tmp/${TAG}_visc python tools/visc_bench.py --n 128 > gpurun_out/${TAG}_ncu3.log 2>&1
python tools/ncu_summary.py /tmp/${TAG}_visc.ncu-rep > gpurun_out/${TAG}_visc_f64.summary.txt 2>&1
du -sh gpurun_out; ls -la gpurun_out | grep ${TAG}_
