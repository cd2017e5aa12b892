#!/usr/bin/env python
"""Steady steps of the sharded 512x512x512N scene under torchrun, for a per-rank ncu launch list (experiment tool)."""
import ctypes
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist

from dcgrid_b200 import FluidSimulationDCGridSharded, scene_params  # noqa: E402

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
d, M = 512, 524288
size = (d, d, d * world)
rt = ctypes.CDLL("libcudart.so")
sim = FluidSimulationDCGridSharded(size, M * world, scene_params(*size, solids=True), world, rank=rank, nlocal=1, device=local, dist=dist)
sim.step(int(os.environ.get("WARM", "30")))
dist.barrier()
rt.cudaProfilerStart()
sim.step(2)
rt.cudaProfilerStop()
print("rank", rank, "steady", bool(sim.counters()[7]), "ms/step", sim.lastStepMs() / 2, flush=True)
dist.barrier()
sim.close()
dist.destroy_process_group()
