#!/bin/bash
# Round 2, GPU pass B: the bench-size golden tests after the FNV-basis fix, then a short bench (parity flag).
set -x
export LD_LIBRARY_PATH=/usr/local/cuda/lib64:$LD_LIBRARY_PATH
TAG=${1:-r2b}
( time python -m pytest tests/test_golden_big_gpu.py -q -m gpu --durations=8 ) > gpurun_out/${TAG}_pytest_big.log 2>&1; tail -15 gpurun_out/${TAG}_pytest_big.log
python bench.py --steps 200 --warmup 10 --no-cpu-baseline --no-reference-cuda > gpurun_out/${TAG}_bench_c3.json 2> gpurun_out/${TAG}_bench_c3.err; python -c "
import json;d=json.load(open('gpurun_out/${TAG}_bench_c3.json'));print(d['ms_per_step'],d['parity'])"
