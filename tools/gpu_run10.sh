set -x
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"k_dc" -c 40000 --csv --log-file gpurun_out/launches_steady.csv python tools/exp_stage.py --warm 130 --steps 4 > gpurun_out/ncu_ls.log 2>&1
tail -2 gpurun_out/ncu_ls.log
wc -l gpurun_out/launches_steady.csv
