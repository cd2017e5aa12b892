set -x
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29551 tests/mgpu_dcgrid_check.py --steps 6 --bench-size 0 > gpurun_out/mgpu2_dcgrid.log 2>&1
grep "^{" gpurun_out/mgpu2_dcgrid.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29553 tests/mgpu_uniform_check.py --bench-size 0 > gpurun_out/mgpu2_uniform.log 2>&1
grep "^{" gpurun_out/mgpu2_uniform.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29555 bench.py --gpus 2 --steps 30 --warmup 5 > gpurun_out/r1e_bench_n2.json 2> gpurun_out/r1e_bench_n2.err
grep "^{" gpurun_out/r1e_bench_n2.json | cut -c1-260
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29557 bench.py --impl reference --gpus 2 --steps 3 --warmup 1 > gpurun_out/r1e_bench_ref_n2.json 2> gpurun_out/r1e_bench_ref_n2.err
grep "^{" gpurun_out/r1e_bench_ref_n2.json | cut -c1-260
