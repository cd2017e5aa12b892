#!/usr/bin/env python
"""Launch list of steady-state steps for `ncu --profile-from-start off` (experiment tool).
usage: ncu --profile-from-start off --metrics gpu__time_duration.sum --cache-control none --clock-control none --csv \
           --log-file out.csv python tools/trace_step.py [--steps 2]"""
import argparse
import ctypes
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from dcgrid_b200 import FluidSimulationDCGrid, scene_params  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--warm", type=int, default=140)
ap.add_argument("--steps", type=int, default=2)
ap.add_argument("--d", type=int, default=512)
ap.add_argument("--M", type=int, default=524288)
a = ap.parse_args()
rt = ctypes.CDLL("libcudart.so")
sim = FluidSimulationDCGrid((a.d,) * 3, a.M, scene_params(a.d, solids=True))
sim.step(a.warm)
rt.cudaProfilerStart()
sim.step(a.steps)
rt.cudaProfilerStop()
print("steady", bool(sim.counters()[7]), "ms/step", sim.lastStepMs() / a.steps)
