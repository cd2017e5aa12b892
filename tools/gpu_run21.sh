set -x
export LD_LIBRARY_PATH=/usr/local/cuda/lib64:$LD_LIBRARY_PATH
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv; nproc; lscpu | grep "Model name"
python -m pytest tests -q -m gpu 2>&1 | tail -4
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
python bench.py --steps 50 --warmup 10 > gpurun_out/r1d_bench_c3.json 2> gpurun_out/r1d_bench_c3.err; tail -2 gpurun_out/r1d_bench_c3.err; cat gpurun_out/r1d_bench_c3.json
python bench.py --workload dcgrid256 --steps 50 --warmup 10 --no-reference-cuda --no-cpu-baseline > gpurun_out/r1d_bench_c2.json 2> gpurun_out/r1d_bench_c2.err; cat gpurun_out/r1d_bench_c2.json
python bench.py --workload uniform64 --steps 100 --warmup 10 --no-reference-cuda > gpurun_out/r1d_bench_c1.json 2> gpurun_out/r1d_bench_c1.err; cat gpurun_out/r1d_bench_c1.json
timeout 600 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:"advect_pipe|divergence_pipe|apply_pipe|prolongate_staged" -c 8 -o gpurun_out/r1d_top_a -f python tools/trace_step.py --steps 1 > gpurun_out/ncu_a.log 2>&1
timeout 600 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:"jacobi_pipe" --launch-skip 10 -c 12 -o gpurun_out/r1d_top_b -f python tools/trace_step.py --steps 1 > gpurun_out/ncu_b.log 2>&1
ncu -i gpurun_out/r1d_top_a.ncu-rep --page raw --csv > gpurun_out/r1d_top_a_raw.csv 2>/dev/null
ncu -i gpurun_out/r1d_top_b.ncu-rep --page raw --csv > gpurun_out/r1d_top_b_raw.csv 2>/dev/null
rm -f gpurun_out/r1d_top_b.ncu-rep
ls -la gpurun_out | tail -12
