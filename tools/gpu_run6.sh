set -x
python -m pytest tests -x -q -m gpu 2>&1 | tail -15
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
timeout 300 compute-sanitizer --tool memcheck --error-exitcode 9 python -c "
from dcgrid_b200 import FluidSimulationDCGrid, scene_params
s = FluidSimulationDCGrid((64,64,64), 2000, scene_params(64, solids=True)); s.step(12); print('sanitizer run ok', s.counters())
" 2>&1 | tail -8
python bench.py --steps 50 --warmup 10 --no-reference-cuda > gpurun_out/bench_c3.json 2> gpurun_out/bench_c3.err; tail -2 gpurun_out/bench_c3.err; cat gpurun_out/bench_c3.json
python bench.py --steps 50 --warmup 130 --no-reference-cuda --no-cpu-baseline > gpurun_out/bench_c3_steady.json 2> gpurun_out/bench_c3_steady.err; tail -2 gpurun_out/bench_c3_steady.err; cat gpurun_out/bench_c3_steady.json
python bench.py --workload dcgrid256 --steps 50 --warmup 10 --no-reference-cuda --no-cpu-baseline > gpurun_out/bench_c2.json 2> gpurun_out/bench_c2.err; tail -2 gpurun_out/bench_c2.err; cat gpurun_out/bench_c2.json
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/launches_r1.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-reference-cuda > gpurun_out/ncu_b.log 2>&1
tail -2 gpurun_out/ncu_b.log
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"k_dc_jacobi4|k_dc_advect_velocity|k_dc_divergence4|k_dc_apply_pressure4|k_dc_advect_density" --launch-skip 95 -c 12 -o gpurun_out/top_r1 -f python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-reference-cuda > gpurun_out/ncu_full.log 2>&1
tail -2 gpurun_out/ncu_full.log
ls -la gpurun_out; du -sm gpurun_out
if [ $(du -sm gpurun_out | cut -f1) -gt 60 ]; then rm -f gpurun_out/top_r1.ncu-rep; fi
