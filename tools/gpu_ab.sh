#!/bin/bash
# A/B of kernel variants: bash tools/gpu_ab.sh "experiment=0" "experiment=1" ...  (each argument = one option set of tools/jacobi_time.py)
export LD_LIBRARY_PATH=/usr/local/cuda/lib64:$LD_LIBRARY_PATH
for o in "$@"; do
  echo "== $o"
  python tools/jacobi_time.py $o 2>&1
done | tee gpurun_out/ab.log
