import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from dcgrid_b200 import FluidSimulationDCGrid, scene_params
d, M, n = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
s = FluidSimulationDCGrid((d, d, d), M, scene_params(d, solids=bool(int(os.environ.get('SOLIDS','1')))))
print("created", flush=True)
for i in range(n):
    s.advectVelocity(); s.synchronize(); print("av", flush=True)
    s.adaptTopology(); s.synchronize(); print("adapt", flush=True)
    s.project(); s.synchronize(); print("project", flush=True)
    s.advectDensity(); print("ad", flush=True)
    s.synchronize(); print("step", i, flush=True)
print("ok", s.totalDensity(), s.counters(), flush=True)
