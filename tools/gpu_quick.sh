#!/bin/bash
# Quick GPU pass for kernel A/B work: `gpurun -- 'bash tools/gpu_quick.sh [tag] [pytest-args]'`:
# the parity tests (all, or the selection given), then digest + step + per-stage times of the dcgrid512 scene.
export LD_LIBRARY_PATH=/usr/local/cuda/lib64:$LD_LIBRARY_PATH
TAG=${1:-quick}
shift
if [ "$1" != "notest" ]; then
  python -m pytest tests -q -x -m gpu "$@" 2>&1 | tail -4 | tee gpurun_out/${TAG}_pytest.log
fi
python tools/jacobi_time.py 2>&1 | tee gpurun_out/${TAG}_stages.log
