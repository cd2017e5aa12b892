set -x
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29519 bench.py --gpus 2 --steps 30 --warmup 5 > gpurun_out/r1d_bench_n2.json 2> gpurun_out/r1d_bench_n2.err
tail -5 gpurun_out/r1d_bench_n2.err; cat gpurun_out/r1d_bench_n2.json
