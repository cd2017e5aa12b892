#!/bin/bash
# Round 2, GPU pass E: whole GPU suite with the device-side selection, transient split, bench.
set -x
export LD_LIBRARY_PATH=/usr/local/cuda/lib64:$LD_LIBRARY_PATH
TAG=${1:-r2e}
( time python -m pytest tests -q -m gpu -x --durations=8 ) > gpurun_out/${TAG}_pytest.log 2>&1; tail -15 gpurun_out/${TAG}_pytest.log
python tools/transient_split.py > gpurun_out/${TAG}_transient_split.log 2>&1; cat gpurun_out/${TAG}_transient_split.log
python bench.py --steps 200 --warmup 10 --no-cpu-baseline --no-reference-cuda > gpurun_out/${TAG}_bench_c3.json 2> gpurun_out/${TAG}_bench_c3.err; python -c "
import json;d=json.load(open('gpurun_out/${TAG}_bench_c3.json'));print(d['ms_per_step'],d['parity']['ok'],d['transient'])"
