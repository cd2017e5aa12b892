set -x
python -m pytest tests -q -m gpu -x 2>&1 | tail -8
python tools/exp_stage.py apply_pressure divergence 2>&1 | tail -1
DCG_COARSE=gmem python tools/exp_stage.py 2>&1 | tail -1
