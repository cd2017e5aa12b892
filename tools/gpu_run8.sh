set -x
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv; nproc; lscpu | grep "Model name"
python -m pytest tests -q -m gpu 2>&1 | tail -15
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
python bench.py --steps 50 --warmup 10 > gpurun_out/bench_c3.json 2> gpurun_out/bench_c3.err; tail -2 gpurun_out/bench_c3.err; cat gpurun_out/bench_c3.json
python bench.py --steps 50 --warmup 130 --no-reference-cuda --no-cpu-baseline > gpurun_out/bench_c3_steady.json 2> gpurun_out/bench_c3_steady.err; tail -2 gpurun_out/bench_c3_steady.err; cat gpurun_out/bench_c3_steady.json
python bench.py --workload dcgrid256 --steps 50 --warmup 10 --no-reference-cuda --no-cpu-baseline > gpurun_out/bench_c2.json 2> gpurun_out/bench_c2.err; cat gpurun_out/bench_c2.json
python bench.py --workload uniform64 --steps 100 --warmup 10 > gpurun_out/bench_c1.json 2> gpurun_out/bench_c1.err; cat gpurun_out/bench_c1.json
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; cat gpurun_out/bench_ref.json
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/launches_r1.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-reference-cuda > gpurun_out/ncu_b.log 2>&1
tail -2 gpurun_out/ncu_b.log
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"k_dc_jacobi4|k_dc_advect_velocity|k_dc_divergence4|k_dc_apply_pressure4|k_dc_advect_density" --launch-skip 104 -c 6 -o gpurun_out/top_r1 -f python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-reference-cuda > gpurun_out/ncu_full.log 2>&1
tail -2 gpurun_out/ncu_full.log
ls -la gpurun_out; du -sm gpurun_out
if [ $(du -sm gpurun_out | cut -f1) -gt 60 ]; then rm -f gpurun_out/top_r1.ncu-rep; fi
