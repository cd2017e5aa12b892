"""The dense solver at sizes that engage its z-marching kernels (x >= 128 cells), against outputs of the REFERENCE's own
CUDA kernels (tests/golden/bigu/*.npz, made by tests/golden/make_golden_uniform_big.py on a B200) — on one GPU and as 2 / 4
ranks of the slab decomposition (all ranks in one process: the same kernels on z-ranges of the same arrays; the
one-rank-per-process path over NVLink is tests/mgpu_uniform_check.py under torchrun).  Bit-exact: FNV-1a-64 of the raw
density + velocity arrays, SHA-256 and a strided sample of density / velocity / pressure / divergence."""
import glob
import json
import os

import numpy as np
import pytest

from dcgrid_b200 import FluidSimulationUniform, FluidSimulationUniformSharded, fnv1a64, scene_params
from tests.test_oracle_golden import sha

pytestmark = pytest.mark.gpu
BIG = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "bigu")
CASES = sorted(os.path.basename(f)[:-4] for f in glob.glob(os.path.join(BIG, "*.npz")))


def test_uniform_fixtures_are_present():
    assert len(CASES) >= 2, "tests/golden/bigu/*.npz missing: run tests/golden/make_golden_uniform_big.py on the GPU box"


@pytest.mark.parametrize("world", [1, 2, 4])
@pytest.mark.parametrize("case", CASES)
def test_uniform_reproduces_reference_cuda(gpu, case, world):
    z = np.load(os.path.join(BIG, case + ".npz"))
    meta = json.loads(bytes(z["meta"]).decode())
    size = (meta["gx"], meta["gy"], meta["gz"])
    if meta["digest_only"] and world == 4:
        pytest.skip("the digest-only case runs on one GPU and as two ranks")
    p = scene_params(*size, solids=bool(meta["solids"]))
    sim = FluidSimulationUniform(size, p) if world == 1 else FluidSimulationUniformSharded(size, p, world)
    local = meta["schedule"] == "local"
    out = {}
    for s in range(meta["steps"]):
        sim.advectVelocity(); sim.adaptTopology()
        sim.projectLocal() if local else sim.project()
        if s == meta["steps"] - 1 and not meta["digest_only"]:
            for f in ("pressure", "divergence"):
                out[f] = sim.field(f).copy()
        sim.advectDensity()
    for f in ("density", "velocity"):
        out[f] = sim.field(f)
    assert fnv1a64(out["density"], out["velocity"]) == int(bytes(z["fnv_raw_density_velocity"]).hex(), 16), "FNV-1a of raw density + velocity"
    if meta["digest_only"]:
        return
    for f in ("density", "velocity", "pressure", "divergence"):
        got = np.ascontiguousarray(out[f], dtype=np.float32)
        np.testing.assert_array_equal(got.reshape(-1)[::meta["sample_stride"]].view(np.uint32), z["sample_" + f].view(np.uint32), err_msg=f"{f} (strided sample)")
        assert sha(got) == bytes(z["sha_" + f]).hex(), f"{f}: SHA-256 of the full array"
