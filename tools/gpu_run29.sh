set -x
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29531 tests/mgpu_dcgrid_check.py --steps 5 --bench-size 0 > gpurun_out/mgpu4_dcgrid.log 2>&1
tail -3 gpurun_out/mgpu4_dcgrid.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 4 --steps 20 --warmup 3 > gpurun_out/r1d_bench_n4.json 2> gpurun_out/r1d_bench_n4.err
tail -3 gpurun_out/r1d_bench_n4.err; cut -c1-700 gpurun_out/r1d_bench_n4.json
