set -x
python -m pytest tests -q -m gpu -x 2>&1 | tail -12
python tools/exp_stage.py divergence 2>&1 | tail -1
