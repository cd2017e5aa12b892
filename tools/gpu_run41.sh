set -x
python -m pytest tests -q -m gpu -x 2>&1 | tail -3
python tools/exp_stage.py jacobi:2 2>&1 | tail -1
python bench.py --workload dcgrid256 --steps 50 --warmup 10 --no-reference-cuda --no-cpu-baseline 2>/dev/null | cut -c1-260
