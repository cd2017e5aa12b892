set -x
python -m pytest tests -q -m gpu -x 2>&1 | tail -8
python tools/exp_stage.py jacobi:0 jacobi:1 advect_both divergence apply_pressure prolongate:0 2>&1 | tail -1
DCG_RESORT=0 python tools/exp_stage.py jacobi:0 jacobi:1 advect_both divergence apply_pressure prolongate:0 2>&1 | tail -1
