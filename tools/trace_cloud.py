#!/usr/bin/env python
"""Steady steps of the cloud scene (terrain + temperature / vapor + fused sources) for an ncu launch list (experiment tool)."""
import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from dcgrid_b200 import FluidSimulationDCGrid, make_ext, scene_params
d, M = 512, 524288
rt = ctypes.CDLL("libcudart.so")
sim = FluidSimulationDCGrid((d, d, d), M, scene_params(d, solids=True))
sim.setExt(make_ext(terrain=1, terrain_height=96.0, terrain_wavelength=128.0, sources=1))
sim.reset()
sim.step(140)
rt.cudaProfilerStart()
sim.step(2)
rt.cudaProfilerStop()
print("steady", bool(sim.counters()[7]), "ms/step", sim.lastStepMs() / 2)
