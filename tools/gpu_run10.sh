set -x
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"k_dc_advect_pipe" --launch-skip 300 -c 2 -o gpurun_out/advp_steady -f python tools/exp_stage.py --reps 4 advect_velocity advect_density > gpurun_out/ncu_adv.log 2>&1
tail -3 gpurun_out/ncu_adv.log
