set -x
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv
nproc; lscpu | grep "Model name"
mkdir -p gpurun_out/golden
python tests/golden/make_golden.py gpurun_out/golden > gpurun_out/golden/log.txt 2>&1; echo "golden rc=$?"
tail -5 gpurun_out/golden/log.txt
python -m pytest tests/test_uniform_gpu.py -x -q -m gpu 2>&1 | tail -15
# reference CUDA timing on bigger configs (no dumps)
oracle/_ref/ref_harness grid=dcgrid d=256 M=65536 solids=0 steps=20 > gpurun_out/ref_c2.json 2>&1
oracle/_ref/ref_harness grid=dcgrid d=512 M=524288 solids=1 steps=20 > gpurun_out/ref_c3.json 2>&1
oracle/_ref/ref_harness grid=uniform d=64 steps=100 > gpurun_out/ref_c1.json 2>&1
cat gpurun_out/ref_c*.json
