import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from dcgrid_b200 import FluidSimulationDCGrid, scene_params
d, M = int(sys.argv[1]), int(sys.argv[2])
s = FluidSimulationDCGrid((d, d, d), M, scene_params(d, solids=False))
print("created", s.levelTable()["loads"], flush=True)
s.advectVelocity(); s.synchronize(); print("av", flush=True)
s.adaptTopology(); s.synchronize(); print("adapt", flush=True)
for st, lvl in (("divergence", 0), ("jacobi", 0), ("jacobi", 1), ("prolongate", 0), ("accumulate_velocity", 0), ("apply_pressure", 0)):
    print("->", st, lvl, flush=True)
    print(s.benchStage(st, lvl, 1), flush=True)
print("done", flush=True)
