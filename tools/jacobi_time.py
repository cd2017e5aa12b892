#!/usr/bin/env python
"""Level-0 Jacobi sweep and step time of the dcgrid512 scene (dcg_bench_stage), for kernel A/B runs."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from dcgrid_b200 import FluidSimulationDCGrid, fnv1a64, scene_params
d, M = 512, 524288
opts = {k: int(v) for k, v in (o.split("=") for o in sys.argv[1:])}
sim = FluidSimulationDCGrid((d, d, d), M, scene_params(d, solids=True), options=opts or None)
sim.step(140)
print("digest", f"{fnv1a64(sim.field('density'), sim.field('velocity')):016x}", "(reference de3c5fe71f620ebe)")
sim.step(10)
sim.step(200)
print("ms/step", sim.lastStepMs() / 200)
for st, lv in (("jacobi", 0), ("jacobi", 1), ("advect_both", 0), ("divergence", 0), ("apply_pressure", 0), ("prolongate", 0)):
    sim.benchStage(st, lv, 10)
    ms, b = sim.benchStage(st, lv, 40)
    print(f"{st:16s} L{lv} {ms * 1e3:8.1f} us  {b / (ms * 1e-3) / 1e9:8.1f} GB/s alg")
