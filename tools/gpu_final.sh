#!/bin/bash
# Round-end pass on one B200: the whole GPU suite, the uniform 1024^3 evidence (bench line, launch list, ncu --set full of the
# z-marching kernels) and the default bench line.  `gpurun -- 'bash tools/gpu_final.sh TAG'`; results in gpurun_out/.
set -x
export LD_LIBRARY_PATH=/usr/local/cuda/lib64:$LD_LIBRARY_PATH
TAG=${1:-run}
python -m pytest tests -q -m gpu 2>&1 | tail -3 | tee gpurun_out/${TAG}_pytest.log
bash tools/gpu_uniform_evidence.sh $TAG
timeout 300 python bench.py --steps 50 --warmup 10 > gpurun_out/${TAG}_bench_c3.json 2> gpurun_out/${TAG}_bench_c3.err; cut -c1-300 gpurun_out/${TAG}_bench_c3.json
