#!/bin/bash
# N-GPU bench A/B: bash tools/gpu_ab_mgpu.sh N "opt-string" "opt-string" ...   ("" = defaults)
N=$1; shift
for o in "$@"; do
  echo "== N=$N opt=$o"
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29581 bench.py --gpus $N --steps 60 --warmup 5 \
    --no-cpu-baseline --no-reference-cuda --no-named-configs --opt "$o" 2>/dev/null | grep "^{" | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('ms/step', round(d['ms_per_step'],4), 'launches/step', d['diagnostics'].get('graph_launches_per_step'), 'barriers', d['diagnostics'].get('barriers'), 'parity', d.get('parity',{}).get('ok'))"
done | tee gpurun_out/ab_mgpu.log
