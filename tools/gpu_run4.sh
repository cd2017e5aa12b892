set -x
python -m pytest tests -x -q -m gpu 2>&1 | tail -15
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
python bench.py --workload dcgrid64 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_dcgrid64.json 2> gpurun_out/bench_dcgrid64.err; tail -2 gpurun_out/bench_dcgrid64.err; cat gpurun_out/bench_dcgrid64.json
python bench.py --workload uniform64 --steps 100 --warmup 10 > gpurun_out/bench_uniform64.json 2> gpurun_out/bench_uniform64.err; tail -2 gpurun_out/bench_uniform64.err; cat gpurun_out/bench_uniform64.json
python bench.py --steps 50 --warmup 10 > gpurun_out/bench_c3.json 2> gpurun_out/bench_c3.err; tail -2 gpurun_out/bench_c3.err; cat gpurun_out/bench_c3.json
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_r1.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-reference-cuda > gpurun_out/ncu_b.log 2>&1
tail -3 gpurun_out/ncu_b.log
