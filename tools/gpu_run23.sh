set -x
nvidia-smi --query-gpu=index,name --format=csv
nvidia-smi topo -m | head -8
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 tests/mgpu_dcgrid_check.py --steps 6 > gpurun_out/mgpu2_dcgrid.log 2>&1
tail -25 gpurun_out/mgpu2_dcgrid.log
