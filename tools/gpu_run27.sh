set -x
python -m pytest tests -q -m gpu -x 2>&1 | tail -6
python tools/exp_stage.py jacobi:0 jacobi:1 jacobi:2 2>&1 | tail -1
DCG_JACOBI=pipe4 python tools/exp_stage.py jacobi:0 jacobi:1 jacobi:2 2>&1 | tail -1
DCG_JACOBI_CTAS=4 python tools/exp_stage.py jacobi:0 jacobi:1 2>&1 | tail -1
DCG_JACOBI_CTAS=3 python tools/exp_stage.py jacobi:0 jacobi:1 2>&1 | tail -1
