set -x
python -m pytest tests -q -m gpu -x 2>&1 | tail -8
python tools/exp_stage.py advect_velocity advect_density advect_both divergence apply_pressure jacobi:0 jacobi:1 prolongate:0 2>&1 | tail -1
DCG_ADVECT_MINB=3 python tools/exp_stage.py advect_velocity advect_density advect_both 2>&1 | tail -1
DCG_ADVECT_FUSE=0 python tools/exp_stage.py 2>&1 | tail -1
DCG_ZERO_ALL=1 python tools/exp_stage.py divergence 2>&1 | tail -1
