set -x
python -m pytest tests -q -m gpu -x 2>&1 | tail -8
python tools/exp_stage.py advect_velocity advect_density advect_both 2>&1 | tail -1
DCG_ADVECT_MINB=3 python tools/exp_stage.py advect_velocity advect_density advect_both 2>&1 | tail -1
DCG_ADVECT_ORDER=slot python tools/exp_stage.py advect_velocity advect_density advect_both 2>&1 | tail -1
DCG_ADVECT_ORDER=slot DCG_ADVECT_MINB=3 python tools/exp_stage.py advect_velocity advect_density advect_both 2>&1 | tail -1
