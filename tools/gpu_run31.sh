set -x
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 8 --steps 20 --warmup 3 > gpurun_out/r1d_bench_n8.json 2> gpurun_out/r1d_bench_n8.err
tail -3 gpurun_out/r1d_bench_n8.err; grep "^{" gpurun_out/r1d_bench_n8.json | cut -c1-400
