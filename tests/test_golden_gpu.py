"""The CUDA path (through the C ABI) against the golden vectors in tests/golden/ — outputs of the REFERENCE's own
CUDA kernels captured on a B200 (tests/golden/README.md).  Same bar as the CPU test of the oracle
(tests/test_oracle_golden.py): bit-exact, full arrays for the tiny cases, SHA-256 digests + strided samples for the
larger ones; DCGrid cases compare the canonical (slot-permutation invariant) block map and the raw level loads."""
import json
import os

import numpy as np
import pytest

from dcgrid_b200 import FluidSimulationDCGrid, FluidSimulationUniform, scene_params
from tests import _canon
from tests.test_oracle_golden import CASES, FIELDS, GOLDEN, sha

pytestmark = pytest.mark.gpu


def run_cuda(meta):
    grid, d, M, solids, steps, schedule = (meta[k] for k in ("grid", "d", "M", "solids", "steps", "schedule"))
    p = scene_params(d, solids=bool(solids))
    sim = FluidSimulationDCGrid((d, d, d), M, p) if grid == "dcgrid" else FluidSimulationUniform((d, d, d), p)
    if schedule.startswith("jacobi"):
        sim.setJacobiSchedule(2, 1, int(schedule[6:]))
    out = {}
    for s in range(steps):
        sim.advectVelocity()
        sim.adaptTopology()
        sim.project() if schedule == "project" else sim.projectLocal()
        if s == steps - 1:
            for f in ("pressure", "t_pressure", "divergence"):
                out[f] = sim.field(f).copy()
        sim.advectDensity()
    for f in ("density", "velocity", "fluidity"):
        out[f] = sim.field(f).copy()
    if grid == "dcgrid":
        out["topo"] = sim.topology()
        out["loads"] = sim.levelTable()["loads"].copy()
    return out


@pytest.mark.parametrize("case", CASES)
def test_cuda_reproduces_reference_cuda_outputs(gpu, case):
    z = np.load(os.path.join(GOLDEN, case + ".npz"))
    meta = json.loads(bytes(z["meta"]).decode())
    out = run_cuda(meta)
    if meta["grid"] == "dcgrid":
        canon = _canon.canonical(out["topo"], {f: out[f] for f in FIELDS})
        np.testing.assert_array_equal(out["loads"], z["loads"])
        for k in ("blocks", "parent", "child", "apron"):
            arr = canon[k].astype(np.int32)
            assert sha(arr) == bytes(z["sha_topo_" + k]).hex(), f"topology {k}"
            if "topo_" + k in z:
                np.testing.assert_array_equal(arr, z["topo_" + k])
        got = {f: np.ascontiguousarray(canon[f], dtype=np.float32) for f in FIELDS}
    else:
        got = {f: np.ascontiguousarray(out[f], dtype=np.float32) for f in FIELDS}
    for f in FIELDS:
        if f in z:
            np.testing.assert_array_equal(got[f].view(np.uint32).ravel(), z[f].view(np.uint32).ravel(), err_msg=f)
        else:
            np.testing.assert_array_equal(got[f].reshape(-1)[::meta["sample_stride"]].view(np.uint32),
                                          z["sample_" + f].view(np.uint32), err_msg=f"{f} (strided sample)")
        assert sha(got[f]) == bytes(z["sha_" + f]).hex(), f"{f}: SHA-256 of the full array"
