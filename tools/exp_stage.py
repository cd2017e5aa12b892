#!/usr/bin/env python
"""A/B timing of single stage kernels on the steady-state C3 scene (experiment tool, not the bench).
usage: python tools/exp_stage.py [--warm 130] [--reps 40] stage[:level] ...   (env DCG_* select variants)"""
import argparse
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from dcgrid_b200 import FluidSimulationDCGrid, scene_params  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--warm", type=int, default=130)
ap.add_argument("--reps", type=int, default=40)
ap.add_argument("--d", type=int, default=512)
ap.add_argument("--M", type=int, default=524288)
ap.add_argument("--steps", type=int, default=20)
ap.add_argument("stages", nargs="*")
a = ap.parse_args()
sim = FluidSimulationDCGrid((a.d,) * 3, a.M, scene_params(a.d, solids=True))
sim.step(a.warm)
sim.step(a.steps)
out = {"env": {k: v for k, v in os.environ.items() if k.startswith("DCG_")}, "ms_per_step": sim.lastStepMs() / a.steps,
       "steady": bool(sim.counters()[7]), "loads": [int(x) for x in sim.levelTable()["loads"]]}
for s in a.stages:
    name, _, lvl = s.partition(":")
    lvl = int(lvl or 0)
    sim.benchStage(name, lvl, 6)
    ms, b = sim.benchStage(name, lvl, a.reps)
    out[s] = {"us": round(ms * 1e3, 2), "GBps": round(b / ms / 1e6, 1)}
print(json.dumps(out), flush=True)
