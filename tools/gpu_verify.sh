#!/bin/bash
# The standard GPU pass of a round, for `gpurun -- 'bash tools/gpu_verify.sh'` (one B200):
# parity tests, smoke, the default bench line, the launch list of two steady steps and an ncu --set full capture of the
# hot kernels.  Everything lands in gpurun_out/; copy what should be judged into profiles/.
set -x
export LD_LIBRARY_PATH=/usr/local/cuda/lib64:$LD_LIBRARY_PATH
TAG=${1:-run}
python -m pytest tests -q -m gpu 2>&1 | tail -3
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
python bench.py --steps 50 --warmup 10 > gpurun_out/${TAG}_bench_c3.json 2> gpurun_out/${TAG}_bench_c3.err; cut -c1-300 gpurun_out/${TAG}_bench_c3.json
timeout 300 ncu --profile-from-start off --metrics gpu__time_duration.sum --cache-control none --clock-control none --csv \
  --log-file gpurun_out/${TAG}_launches_steady_step.csv python tools/trace_step.py --steps 2 > gpurun_out/trace.log 2>&1
timeout 600 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:"advect_pipe|divergence_pipe|apply_pipe|prolongate_staged" \
  -c 8 -o gpurun_out/${TAG}_top_a -f python tools/trace_step.py --steps 1 > gpurun_out/ncu_a.log 2>&1
timeout 600 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:"jacobi_pipe" --launch-skip 0 -c 12 \
  -o gpurun_out/${TAG}_top_b -f python tools/trace_step.py --steps 1 > gpurun_out/ncu_b.log 2>&1
ncu -i gpurun_out/${TAG}_top_a.ncu-rep --page raw --csv > gpurun_out/${TAG}_top_a_raw.csv 2>/dev/null
ncu -i gpurun_out/${TAG}_top_b.ncu-rep --page raw --csv > gpurun_out/${TAG}_top_b_raw.csv 2>/dev/null
rm -f gpurun_out/${TAG}_top_a.ncu-rep gpurun_out/${TAG}_top_b.ncu-rep
