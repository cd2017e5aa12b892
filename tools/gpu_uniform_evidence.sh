#!/bin/bash
# Evidence for the uniform solver's final (z-marching) kernels at BASELINE configs[4] size, for
# `gpurun -- 'bash tools/gpu_uniform_evidence.sh TAG'` (one B200): the bench line, the launch list of two steps and an
# ncu --set full capture of every k_u_*_zm launch of one step.  Everything lands in gpurun_out/.
set -x
export LD_LIBRARY_PATH=/usr/local/cuda/lib64:$LD_LIBRARY_PATH
TAG=${1:-run}
timeout 300 python bench.py --workload uniform1024 --steps 20 --warmup 5 --no-cpu-baseline --no-reference-cuda \
  > gpurun_out/${TAG}_bench_u1024.json 2> gpurun_out/${TAG}_bench_u1024.err; cut -c1-400 gpurun_out/${TAG}_bench_u1024.json
timeout 240 ncu --profile-from-start off --metrics gpu__time_duration.sum --cache-control none --clock-control none --csv \
  --log-file gpurun_out/${TAG}_uniform1024_launches.csv python tools/trace_uniform.py 1024 > gpurun_out/trace_u.log 2>&1
timeout 420 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:"_zm" -c 16 \
  -o gpurun_out/${TAG}_uniform_zm -f python tools/trace_uniform.py 1024 > gpurun_out/ncu_u.log 2>&1
ncu -i gpurun_out/${TAG}_uniform_zm.ncu-rep --page raw --csv > gpurun_out/${TAG}_uniform_zm_raw.csv 2>/dev/null
ls -la gpurun_out/ | tail -12
rm -f gpurun_out/${TAG}_uniform_zm.ncu-rep
