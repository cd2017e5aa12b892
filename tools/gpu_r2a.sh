#!/bin/bash
# Round 2, GPU pass A: golden fixtures at the bench sizes, the whole GPU test suite, bench A/B (PDL on/off), launch list and
# ncu --set full of the hot kernels.
set -x
export LD_LIBRARY_PATH=/usr/local/cuda/lib64:$LD_LIBRARY_PATH
TAG=${1:-r2a}
mkdir -p gpurun_out/golden_big
nvidia-smi --query-gpu=name,memory.total --format=csv; free -g | head -2; nproc
( time python tests/golden/make_golden_big.py gpurun_out/golden_big ) > gpurun_out/${TAG}_golden_big.log 2>&1
tail -3 gpurun_out/${TAG}_golden_big.log
mkdir -p tests/golden/big && cp gpurun_out/golden_big/*.npz tests/golden/big/
( time python -m pytest tests -q -m gpu -x --durations=15 ) > gpurun_out/${TAG}_pytest.log 2>&1; tail -25 gpurun_out/${TAG}_pytest.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
python bench.py --steps 200 --warmup 10 > gpurun_out/${TAG}_bench_c3.json 2> gpurun_out/${TAG}_bench_c3.err; cut -c1-400 gpurun_out/${TAG}_bench_c3.json
python bench.py --steps 200 --warmup 10 --opt no_pdl=1 --no-cpu-baseline --no-reference-cuda > gpurun_out/${TAG}_bench_c3_nopdl.json 2> gpurun_out/${TAG}_bench_c3_nopdl.err; cut -c1-300 gpurun_out/${TAG}_bench_c3_nopdl.json
python tools/csrc_digest.py > gpurun_out/${TAG}_csrc_digest.txt
timeout 300 ncu --profile-from-start off --metrics gpu__time_duration.sum --cache-control none --clock-control none --csv \
  --log-file gpurun_out/${TAG}_launches_steady_step.csv python tools/trace_step.py --steps 2 > gpurun_out/trace.log 2>&1
timeout 900 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:"advect_pipe|divergence_pipe|apply_pipe|prolongate_staged" \
  -c 8 -o gpurun_out/${TAG}_top_a -f python tools/trace_step.py --steps 1 > gpurun_out/ncu_a.log 2>&1
timeout 900 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:"jacobi_pipe" --launch-skip 9 -c 2 \
  -o gpurun_out/${TAG}_top_b -f python tools/trace_step.py --steps 1 > gpurun_out/ncu_b.log 2>&1
ncu -i gpurun_out/${TAG}_top_a.ncu-rep --page raw --csv > gpurun_out/${TAG}_top_a_raw.csv 2>/dev/null
ncu -i gpurun_out/${TAG}_top_b.ncu-rep --page raw --csv > gpurun_out/${TAG}_top_b_raw.csv 2>/dev/null
ncu -i gpurun_out/${TAG}_top_a.ncu-rep --page source --csv > gpurun_out/${TAG}_top_a_source.csv 2>/dev/null
ncu -i gpurun_out/${TAG}_top_b.ncu-rep --page source --csv > gpurun_out/${TAG}_top_b_source.csv 2>/dev/null
ls -la gpurun_out/*.ncu-rep gpurun_out/${TAG}_*
# gpurun_out travels back only up to 64 MiB: the CSV exports carry everything read afterwards
rm -f gpurun_out/${TAG}_top_a.ncu-rep gpurun_out/${TAG}_top_b.ncu-rep
du -sh gpurun_out
