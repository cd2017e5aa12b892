set -x
python -m pytest tests -q -m gpu -x 2>&1 | tail -6
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
python bench.py --steps 50 --warmup 130 --no-reference-cuda --no-cpu-baseline > gpurun_out/r1c_bench_c3_steady.json 2> gpurun_out/r1c_bench_c3_steady.err; tail -2 gpurun_out/r1c_bench_c3_steady.err; cat gpurun_out/r1c_bench_c3_steady.json
