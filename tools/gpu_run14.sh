set -x
python -m pytest tests -q -m gpu -x -k "speculative" 2>&1 | tail -4
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"k_dc_advect_pipe" --launch-skip 140 -c 1 -o gpurun_out/adv2 -f python tools/exp_stage.py --reps 2 > gpurun_out/ncu_adv2.log 2>&1
tail -3 gpurun_out/ncu_adv2.log
ncu -i gpurun_out/adv2.ncu-rep --page raw --csv > gpurun_out/adv2_raw.csv 2>/dev/null
ncu -i gpurun_out/adv2.ncu-rep --page source --csv > gpurun_out/adv2_source.csv 2>/dev/null
ls -la gpurun_out/
