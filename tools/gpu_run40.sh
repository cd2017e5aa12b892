set -x
python -m pytest tests -q -m gpu 2>&1 | tail -3
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
python bench.py --steps 50 --warmup 10 --no-reference-cuda > gpurun_out/r1f_bench_c3.json 2> gpurun_out/r1f_bench_c3.err; tail -2 gpurun_out/r1f_bench_c3.err; cut -c1-300 gpurun_out/r1f_bench_c3.json
