set -x
python -m pytest tests -q -m gpu -x -k "non_cubic" 2>&1 | tail -8
