#!/bin/bash
# Round 2, GPU pass C: everything after the proxy-fence fix — whole GPU suite, bench (parity flag), launch list.
set -x
export LD_LIBRARY_PATH=/usr/local/cuda/lib64:$LD_LIBRARY_PATH
TAG=${1:-r2c}
( time python -m pytest tests -q -m gpu --durations=8 ) > gpurun_out/${TAG}_pytest.log 2>&1; tail -15 gpurun_out/${TAG}_pytest.log
python bench.py --steps 200 --warmup 10 > gpurun_out/${TAG}_bench_c3.json 2> gpurun_out/${TAG}_bench_c3.err; python -c "
import json;d=json.load(open('gpurun_out/${TAG}_bench_c3.json'));print(d['ms_per_step'],d['parity'],d['stages'],d['roofline']['frac'],d['transient']['ms_per_step'])"
timeout 300 ncu --profile-from-start off --metrics gpu__time_duration.sum --cache-control none --clock-control none --csv \
  --log-file gpurun_out/${TAG}_launches_steady_step.csv python tools/trace_step.py --steps 2 > gpurun_out/trace.log 2>&1
python tools/sum_launches.py gpurun_out/${TAG}_launches_steady_step.csv | tail -25
