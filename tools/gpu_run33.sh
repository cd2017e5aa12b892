set -x
python tools/exp_stage.py --warm 0 --steps 20 2>&1 | tail -1
DCG_RESORT=0 python tools/exp_stage.py --warm 0 --steps 20 2>&1 | tail -1
DCG_JACOBI=pipe4 python tools/exp_stage.py --warm 0 --steps 20 2>&1 | tail -1
DCG_RESORT_EVERY=8 python tools/exp_stage.py --warm 0 --steps 20 2>&1 | tail -1
