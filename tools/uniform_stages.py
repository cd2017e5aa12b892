#!/usr/bin/env python
"""Per-stage device time of the uniform solver (dcg_bench_stage) against the algorithmic bytes of SURVEY §8(d)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from dcgrid_b200 import FluidSimulationUniform, scene_params
d = int(sys.argv[1]) if len(sys.argv) > 1 else 512
opts = {k: int(v) for k, v in (o.split("=") for o in sys.argv[2:])}
sim = FluidSimulationUniform((d, d, d), scene_params(d), options=opts or None)
sim.step(30)
print("ms/step", sim.lastStepMs() / 30, "alg GB/step", sim.algorithmicBytes()[0] / 1e9, "=> GB/s", sim.algorithmicBytes()[0] / (sim.lastStepMs() / 30 * 1e-3) / 1e9)
for st, lv in (("advect_both", 0), ("advect_velocity", 0), ("advect_density", 0), ("divergence", 0), ("jacobi", 0), ("jacobi", 1), ("jacobi", 2), ("apply_pressure", 0)):
    ms, b = sim.benchStage(st, lv, 10)
    print(f"{st:16s} L{lv} {ms:8.3f} ms  {b / (ms * 1e-3) / 1e9:8.1f} GB/s alg")
