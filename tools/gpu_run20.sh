set -x
python -m pytest tests -q -m gpu -x 2>&1 | tail -8
python tools/exp_stage.py prolongate:0 prolongate:1 2>&1 | tail -1
export LD_LIBRARY_PATH=/usr/local/cuda/lib64:$LD_LIBRARY_PATH
timeout 300 ncu --profile-from-start off --metrics gpu__time_duration.sum --cache-control none --clock-control none --csv --log-file gpurun_out/r1d_launches_steady_step.csv python tools/trace_step.py --steps 2 > gpurun_out/trace.log 2>&1
tail -2 gpurun_out/trace.log
