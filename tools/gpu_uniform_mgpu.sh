#!/bin/bash
# Uniform solver on N GPUs: multi-process parity against 1 GPU, then BASELINE configs[4]'s uniform 1024^3 sharded over the N GPUs.
# `gpurun --gpus N -- 'bash tools/gpu_uniform_mgpu.sh N TAG'`
set -x
N=${1:-2}; TAG=${2:-run}
timeout 150 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29563 tests/mgpu_uniform_check.py --bench-size 0 > gpurun_out/${TAG}_mgpu${N}_uniform.log 2>&1
grep "^{" gpurun_out/${TAG}_mgpu${N}_uniform.log | cut -c1-400
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29591 bench.py --gpus $N --workload uniform1024 --strong \
  --steps 10 --warmup 3 --no-cpu-baseline --no-reference-cuda --no-named-configs > gpurun_out/${TAG}_bench_c5_uniform1024_n${N}.json 2> gpurun_out/${TAG}_bench_c5_uniform1024_n${N}.err
grep "^{" gpurun_out/${TAG}_bench_c5_uniform1024_n${N}.json | cut -c1-300
