set -x
DCG_ADVECT_MINB=2 python tools/exp_stage.py advect_both 2>&1 | tail -1
python tools/exp_stage.py advect_both 2>&1 | tail -1
