set -x
python -m pytest tests/test_dcgrid_gpu.py -q -m gpu -x 2>&1 | tail -5
python tools/exp_stage.py jacobi_legacy:0 jacobi_pipe:0 jacobi_legacy:1 jacobi_pipe:1 jacobi_legacy:2 jacobi_pipe:2
DCG_SNAKE=0 python tools/exp_stage.py jacobi_pipe:0 jacobi_pipe:1
DCG_JACOBI_CTAS=3 python tools/exp_stage.py jacobi_pipe:0 jacobi_pipe:1
