set -x
export DCG_ADVECT_MINB=3
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"k_dc_jacobi_pipe|k_dc_divergence4|k_dc_apply_pressure4|k_dc_prolongate4|k_dc_accumulate" --launch-skip 5000 -c 60 -o gpurun_out/sten -f python tools/exp_stage.py --reps 2 > gpurun_out/ncu_sten.log 2>&1
tail -3 gpurun_out/ncu_sten.log
ncu -i gpurun_out/sten.ncu-rep --page raw --csv > gpurun_out/sten_raw.csv 2>/dev/null
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"k_dc_advect_pipe" --launch-skip 140 -c 1 -o gpurun_out/adv3 -f python tools/exp_stage.py --reps 2 > gpurun_out/ncu_adv3.log 2>&1
ncu -i gpurun_out/adv3.ncu-rep --page raw --csv > gpurun_out/adv3_raw.csv 2>/dev/null
ncu -i gpurun_out/adv3.ncu-rep --page source --csv > gpurun_out/adv3_source.csv 2>/dev/null
rm -f gpurun_out/sten.ncu-rep
ls -la gpurun_out/
