#!/bin/bash
# Bench lines of the small BASELINE configs (configs[0] uniform 64^3, configs[1] DCGrid 256^3) at HEAD, then the default line.
# `gpurun -- 'bash tools/gpu_configs.sh TAG'`, one B200; results in gpurun_out/.
set -x
export LD_LIBRARY_PATH=/usr/local/cuda/lib64:$LD_LIBRARY_PATH
TAG=${1:-run}
timeout 200 python bench.py --workload dcgrid256 --steps 200 --warmup 10 --no-named-configs > gpurun_out/${TAG}_bench_c2.json 2> gpurun_out/${TAG}_bench_c2.err; cut -c1-300 gpurun_out/${TAG}_bench_c2.json
timeout 200 python bench.py --workload uniform64 --steps 200 --warmup 10 > gpurun_out/${TAG}_bench_c1.json 2> gpurun_out/${TAG}_bench_c1.err; cut -c1-300 gpurun_out/${TAG}_bench_c1.json
timeout 300 python bench.py > gpurun_out/${TAG}_bench_c3.json 2> gpurun_out/${TAG}_bench_c3.err; cut -c1-300 gpurun_out/${TAG}_bench_c3.json
