#!/bin/bash
# BASELINE configs[4] on N GPUs (strong scaling of the named scenes): bash tools/gpu_c5.sh N TAG
N=$1; TAG=${2:-run}
for wl in dcgrid2048 uniform1024; do
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29591 bench.py --gpus $N --workload $wl --strong \
    --steps 10 --warmup 3 --no-cpu-baseline --no-reference-cuda --no-named-configs > gpurun_out/${TAG}_bench_c5_${wl}_n${N}.json 2> gpurun_out/${TAG}_bench_c5_${wl}_n${N}.err
  grep "^{" gpurun_out/${TAG}_bench_c5_${wl}_n${N}.json | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('$wl N=$N ms/step', round(d['ms_per_step'],3), 'value', '%.3e' % d['value'], 'parity', d.get('parity',{}).get('ok'), d['clocks'].get('reasons'))"
done
