#!/usr/bin/env python
"""A few steps of the uniform solver for an ncu capture (profile-from-start off)."""
import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from dcgrid_b200 import FluidSimulationUniform, scene_params
d = int(sys.argv[1]) if len(sys.argv) > 1 else 512
rt = ctypes.CDLL("libcudart.so")
sim = FluidSimulationUniform((d, d, d), scene_params(d), options={"no_pdl": 1})
sim.step(20)
rt.cudaProfilerStart()
for _ in range(2):
    sim.advectVelocity(); sim.adaptTopology(); sim.project(); sim.advectDensity()
sim.synchronize()
rt.cudaProfilerStop()
