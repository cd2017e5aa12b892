#!/bin/bash
# torchrun entry: rank 0 runs under ncu (launch list), the other ranks run plain
export LD_LIBRARY_PATH=/usr/local/cuda/lib64:$LD_LIBRARY_PATH
if [ "$LOCAL_RANK" = "0" ]; then
  exec ncu --profile-from-start off --metrics gpu__time_duration.sum --cache-control none --clock-control none --csv --log-file gpurun_out/${TAG:-run}_launches_n${WORLD_SIZE}_rank0.csv python tools/trace_step_mgpu.py
else
  exec python tools/trace_step_mgpu.py
fi
