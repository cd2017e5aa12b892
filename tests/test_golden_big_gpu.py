"""The CUDA path at the sizes bench.py measures, against outputs of the REFERENCE's own CUDA kernels at those sizes
(tests/golden/big/*.npz, made by tests/golden/make_golden_big.py on a B200): BASELINE.json configs[1] (256^3,
M = 65,536, 100 steps), configs[2] (512^3, M = 524,288, solids; 140 steps = bench.py's pre-roll, i.e. through the
adaptation transient, the proven fixed point, resort() and the CUDA-graph replay) and the multi-GPU bench scenes.
Bit-exact: SHA-256 of every canonical field and of the canonical block map, SHA-256 / FNV-1a of the raw (slot-order)
density and velocity arrays, strided samples, raw level loads.  The steps run through dcg_step — the call bench.py
times — except the last one, which is issued call by call to read the pressure fields between project() and
advectDensity()."""
import glob
import json
import os

import numpy as np
import pytest

from dcgrid_b200 import FluidSimulationDCGrid, fnv1a64, scene_params
from tests import _canon
from tests.test_oracle_golden import FIELDS, sha

pytestmark = pytest.mark.gpu
BIG = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "big")
CASES = sorted(os.path.basename(f)[:-4] for f in glob.glob(os.path.join(BIG, "*.npz")))


def test_big_fixtures_are_present():
    assert len(CASES) >= 3, "tests/golden/big/*.npz missing: run tests/golden/make_golden_big.py on the GPU box"


@pytest.mark.parametrize("case", CASES)
def test_cuda_reproduces_reference_cuda_at_bench_sizes(gpu, case):
    z = np.load(os.path.join(BIG, case + ".npz"))
    meta = json.loads(bytes(z["meta"]).decode())
    size = (meta["gx"], meta["gy"], meta["gz"])
    p = scene_params(*size, solids=bool(meta["solids"]))
    sim = FluidSimulationDCGrid(size, meta["M"], p)
    sim.step(meta["steps"] - 1)
    out = {}
    sim.advectVelocity(); sim.adaptTopology(); sim.project()
    if not meta["digest_only"]:
        for f in ("pressure", "t_pressure", "divergence"):
            out[f] = sim.field(f).copy()
    sim.advectDensity()
    for f in ("density", "velocity"):
        out[f] = sim.field(f)
    assert fnv1a64(out["density"], out["velocity"]) == int(bytes(z["fnv_raw_density_velocity"]).hex(), 16), "FNV-1a of raw density + velocity"
    if meta["digest_only"]:
        return
    for f in ("density", "velocity"):
        assert sha(out[f]) == bytes(z["sha_raw_" + f]).hex(), f"raw {f}"
    out["fluidity"] = sim.field("fluidity")
    np.testing.assert_array_equal(sim.levelTable()["loads"], z["loads"])
    canon = _canon.canonical(sim.topology(), {f: out[f] for f in FIELDS})
    for k in ("blocks", "parent", "child", "apron"):
        assert sha(canon[k].astype(np.int32)) == bytes(z["sha_topo_" + k]).hex(), f"topology {k}"
    for f in FIELDS:
        got = np.ascontiguousarray(canon[f], dtype=np.float32)
        np.testing.assert_array_equal(got.reshape(-1)[::meta["sample_stride"]].view(np.uint32), z["sample_" + f].view(np.uint32),
                                      err_msg=f"{f} (strided sample)")
        assert sha(got) == bytes(z["sha_" + f]).hex(), f"{f}: SHA-256 of the full array"
    c = sim.counters()
    if meta["steps"] >= 140:
        assert c[7] == 1 and c[4] > 0, "the run must have reached the fixed point and replayed the step graph"
