set -x
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29521 --no-python tools/ncu_rank0.sh > gpurun_out/trace_n2.log 2>&1
tail -5 gpurun_out/trace_n2.log
