set -x
python -m pytest tests -q -m gpu -x 2>&1 | tail -3
timeout 900 python bench.py --workload dcgrid2048 --steps 10 --warmup 3 --preroll 40 --no-reference-cuda --no-cpu-baseline > gpurun_out/r1d_bench_c5d.json 2> gpurun_out/r1d_bench_c5d.err; tail -3 gpurun_out/r1d_bench_c5d.err; cut -c1-1800 gpurun_out/r1d_bench_c5d.json
timeout 600 python bench.py --workload uniform1024 --steps 10 --warmup 3 --preroll 0 --no-reference-cuda --no-cpu-baseline > gpurun_out/r1d_bench_c5u.json 2> gpurun_out/r1d_bench_c5u.err; tail -3 gpurun_out/r1d_bench_c5u.err; cut -c1-1500 gpurun_out/r1d_bench_c5u.json
nvidia-smi --query-gpu=memory.used --format=csv
