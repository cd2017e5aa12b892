set -x
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29561 tests/mgpu_dcgrid_check.py --steps 6 --bench-size 0 > gpurun_out/mgpu2_dcgrid.log 2>&1
grep "^{" gpurun_out/mgpu2_dcgrid.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29563 bench.py --gpus 2 --steps 30 --warmup 5 > gpurun_out/r1e_bench_n2.json 2> gpurun_out/r1e_bench_n2.err
grep "^{" gpurun_out/r1e_bench_n2.json | cut -c1-260
