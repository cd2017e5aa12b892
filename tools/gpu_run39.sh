set -x
export LD_LIBRARY_PATH=/usr/local/cuda/lib64:$LD_LIBRARY_PATH
timeout 600 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:"jacobi_pipe8" --launch-skip 14 -c 1 -o gpurun_out/j8 -f python tools/trace_step.py --steps 1 > gpurun_out/ncu_j8.log 2>&1
ncu -i gpurun_out/j8.ncu-rep --page source --csv > gpurun_out/j8_source.csv 2>/dev/null
rm -f gpurun_out/j8.ncu-rep
python tools/exp_stage.py jacobi_legacy:2 jacobi:2 jacobi:3 2>&1 | tail -1
