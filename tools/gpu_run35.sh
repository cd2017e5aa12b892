set -x
python -m pytest tests -q -m gpu -x -k "variants or bit_exact_vs_oracle or sharded" 2>&1 | tail -3
python tools/exp_stage.py jacobi:0 jacobi:1 jacobi:2 2>&1 | tail -1
