set -x
python -m pytest tests -q -m gpu -x 2>&1 | tail -4
python tools/exp_stage.py --warm 0 --steps 20 2>&1 | tail -1
python tools/exp_stage.py --warm 20 --steps 60 2>&1 | tail -1
DCG_RESORT=0 python tools/exp_stage.py --warm 20 --steps 60 2>&1 | tail -1
