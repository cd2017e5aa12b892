"""The product driven through the reference's own abstract class: oracle/_ref/adapter_demo is built from
include/fluid_simulation_b200.h + the reference's fluid_simulation.cpp (compiled where it lies) and issues
the call sequence of Simulation::updateSimulation (src/simulation.cpp:93-116) through a FluidSimulation*."""
import json
import os
import subprocess

import numpy as np
import pytest

from dcgrid_b200 import FluidSimulationDCGrid, FluidSimulationUniform, scene_params
from tests._oracle import Oracle

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EXE = os.path.join(ROOT, "oracle", "_ref", "adapter_demo")


def fnv(arrays):
    h = 1469598103934665603
    for a in arrays:
        for b in np.ascontiguousarray(a, dtype=np.float32).tobytes():
            h = ((h ^ b) * 1099511628211) & 0xFFFFFFFFFFFFFFFF
    return f"{h:016x}"


@pytest.mark.parametrize("grid,d,M,solids,steps", [("dcgrid", 32, 300, 1, 6), ("uniform", 32, 0, 0, 5)])
def test_adapter_matches_oracle(gpu, grid, d, M, solids, steps):
    if not os.path.exists(EXE):
        pytest.fail("oracle/_ref/adapter_demo missing: run __graft_entry__.build() where /root/reference is mounted")
    r = subprocess.run([EXE, f"grid={grid}", f"d={d}", f"M={max(M, 1)}", f"solids={solids}", f"steps={steps}"], capture_output=True,
                       text=True, timeout=300)
    assert r.returncode == 0, r.stderr
    got = json.loads([l for l in r.stdout.splitlines() if l.startswith("{")][-1])
    p = scene_params(d, solids=bool(solids))
    orc = Oracle(p, M if grid == "dcgrid" else 0)
    orc.step(steps)
    q, v = orc.field("density"), orc.field("velocity")
    if grid == "dcgrid":
        # free pool slots hold zeros on both sides (memset at reset)
        sim = FluidSimulationDCGrid((d, d, d), M, p)
        sim.step(steps)
        assert got["digest"] == fnv([sim.field("density"), sim.field("velocity")])
        act = np.repeat(orc.topology(with_apron=False)["level"] != 0xFF, 64)
        np.testing.assert_array_equal(sim.field("density")[act].view(np.uint32), q[act].view(np.uint32))
    else:
        assert got["digest"] == fnv([q, v])
