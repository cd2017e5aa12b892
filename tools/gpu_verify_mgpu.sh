#!/bin/bash
# Multi-GPU pass, for `gpurun --gpus N -- 'bash tools/gpu_verify_mgpu.sh N'`: N-rank parity against 1 GPU for both
# solvers, then the bench line of the slab-decomposed scene.
set -x
N=${1:-2}
TAG=${2:-run}
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29561 tests/mgpu_dcgrid_check.py --steps 6 --bench-size 0 > gpurun_out/${TAG}_mgpu${N}_dcgrid.log 2>&1
grep "^{" gpurun_out/${TAG}_mgpu${N}_dcgrid.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29563 tests/mgpu_uniform_check.py --bench-size 0 > gpurun_out/${TAG}_mgpu${N}_uniform.log 2>&1
grep "^{" gpurun_out/${TAG}_mgpu${N}_uniform.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29565 bench.py --gpus $N --steps 30 --warmup 5 > gpurun_out/${TAG}_bench_n${N}.json 2> gpurun_out/${TAG}_bench_n${N}.err
grep "^{" gpurun_out/${TAG}_bench_n${N}.json | cut -c1-300
