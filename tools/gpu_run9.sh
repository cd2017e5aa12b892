set -x
timeout 300 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
timeout 120 python tools/exp_stage.py advect_velocity advect_density divergence apply_pressure
