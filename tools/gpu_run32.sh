set -x
export LD_LIBRARY_PATH=/usr/local/cuda/lib64:$LD_LIBRARY_PATH
python -m pytest tests -q -m gpu 2>&1 | tail -3
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
python bench.py --steps 50 --warmup 10 > gpurun_out/r1e_bench_c3.json 2> gpurun_out/r1e_bench_c3.err; tail -2 gpurun_out/r1e_bench_c3.err; cut -c1-300 gpurun_out/r1e_bench_c3.json
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r1e_bench_ref.json 2> gpurun_out/r1e_bench_ref.err; cut -c1-300 gpurun_out/r1e_bench_ref.json
timeout 300 ncu --profile-from-start off --metrics gpu__time_duration.sum --cache-control none --clock-control none --csv --log-file gpurun_out/r1e_launches_steady_step.csv python tools/trace_step.py --steps 2 > gpurun_out/trace.log 2>&1
tail -1 gpurun_out/trace.log
timeout 600 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:"advect_pipe|divergence_pipe|apply_pipe|prolongate_staged" -c 8 -o gpurun_out/r1e_top_a -f python tools/trace_step.py --steps 1 > gpurun_out/ncu_a.log 2>&1
timeout 600 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:"jacobi_pipe" --launch-skip 10 -c 12 -o gpurun_out/r1e_top_b -f python tools/trace_step.py --steps 1 > gpurun_out/ncu_b.log 2>&1
ncu -i gpurun_out/r1e_top_a.ncu-rep --page raw --csv > gpurun_out/r1e_top_a_raw.csv 2>/dev/null
ncu -i gpurun_out/r1e_top_b.ncu-rep --page raw --csv > gpurun_out/r1e_top_b_raw.csv 2>/dev/null
rm -f gpurun_out/r1e_top_b.ncu-rep gpurun_out/r1e_top_a.ncu-rep
