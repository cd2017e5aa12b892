#!/usr/bin/env python
"""Which dcg_options variants reproduce the reference's own CUDA kernels after `steps` steps of a d^3 scene with solids:
    python tools/jacobi_race_probe.py 512 524288 1 [k=v,k=v ...]
(r2b: before the generic->async proxy fence in the ring kernels every variant that used them differed from the reference
at the FIRST step of the 512^3 scene, with a different digest on every run; profiles/README.md.)"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from dcgrid_b200 import FluidSimulationDCGrid, fnv1a64, scene_params  # noqa: E402
from tests import _refio  # noqa: E402

d, M, steps = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
kw = dict(grid="dcgrid", gx=d, gy=d, gz=d, M=M, solids=1, schedule="project", serialize=1)
lines, _, _ = _refio.run_harness(fma=False, steps=steps, timeout=3000, **kw)
ref = lines[0]["final_digest"]
p = scene_params(d, solids=True)
variants = [{}, {}, {"jacobi": 1}, {"jacobi": 3}, {"no_snake": 1}, {"no_pdl": 1}, {"no_pdl": 1, "no_snake": 1}, {"jacobi_ctas_per_sm": 1}, {"jacobi_ctas_per_sm": 2},
            {"jacobi_ctas_per_sm": 4}, {"jacobi": 2}, {"jacobi": 2, "no_pdl": 1}, {"no_resort": 1}, {"jacobi_max_ctas": 148}]
for v in sys.argv[4:]:
    variants.append({k: int(x) for k, x in (o.split("=") for o in v.split(","))})
for o in variants:
    try:
        sim = FluidSimulationDCGrid((d, d, d), M, p, options=o or None)
    except KeyError as e:
        print("skip", o, e)
        continue
    sim.step(steps)
    got = f"{fnv1a64(sim.field('density'), sim.field('velocity')):016x}"
    print("variant", o, got, "OK" if got == ref else "MISMATCH", flush=True)
    del sim
