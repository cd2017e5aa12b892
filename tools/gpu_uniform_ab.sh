#!/bin/bash
# Uniform solver: parity tests of the z-marching kernels, then the 1024^3 bench line with the shipped kernels and with
# each A/B variant (dcg_options.experiment bits, see uniform.cu).  `gpurun -- 'bash tools/gpu_uniform_ab.sh TAG "0 64 128 256"'`.
set -x
export LD_LIBRARY_PATH=/usr/local/cuda/lib64:$LD_LIBRARY_PATH
TAG=${1:-run}
VARIANTS=${2:-"0"}
python -m pytest tests/test_uniform_gpu.py tests/test_golden_uniform_big_gpu.py tests/test_sharded_gpu.py tests/test_golden_gpu.py -q -m gpu -x 2>&1 | tail -4
for e in $VARIANTS; do
  timeout 200 python bench.py --workload uniform1024 --steps 20 --warmup 5 --no-cpu-baseline --no-reference-cuda --opt experiment=$e \
    > gpurun_out/${TAG}_bench_u1024_e$e.json 2> gpurun_out/${TAG}_bench_u1024_e$e.err
  python - <<PY
import json
d = json.load(open("gpurun_out/${TAG}_bench_u1024_e$e.json"))
print("experiment=$e", round(d["ms_per_step"], 3), "ms/step", {k: round(v["ms"], 3) for k, v in d.get("stages", {}).items()}, "step frac", round(d["step_roofline"]["frac"], 4))
PY
done
