"""Slab decomposition of the adaptive solver: an N-rank run must equal the 1-GPU run — block pool and every
field, bit for bit (same per-cell arithmetic, only the rank that runs a tile differs).  Here all ranks live in
one process on one device (nlocal == world), which exercises the ownership table, the per-rank tile runs, the
per-rank Morton / parent lists and the rank-0-only coarse kernels of csrc/dcgrid.cu; the multi-process path
(one virtual range stitched from every GPU's arena, flag barriers over NVLink) is tests/mgpu_dcgrid_check.py
(torchrun, >= 2 GPUs)."""
import numpy as np
import pytest

from dcgrid_b200 import DcgError, FluidSimulationDCGrid, FluidSimulationDCGridSharded, scene_params
from tests._oracle import Oracle

pytestmark = pytest.mark.gpu


def _bits(a):
    return np.ascontiguousarray(a, dtype=np.float32).view(np.uint32)


def _same_state(a, b, fields, where):
    ta, tb = a.topology(), b.topology()
    for k in ("level", "pos", "parent", "child", "apron"):
        np.testing.assert_array_equal(ta[k], tb[k], err_msg=f"{k} {where}")
    for f in fields:
        np.testing.assert_array_equal(_bits(a.field(f)), _bits(b.field(f)), err_msg=f"{f} {where}")


@pytest.mark.parametrize("d,M,world,solids,steps,unit", [
    (64, 2000, 2, True, 10, None),      # one-tile ownership granularity: ranks interleave inside every level
    (64, 2000, 3, False, 8, "64"),
    (64, 4096, 4, False, 12, "128"),    # perpetual move cycle: the per-rank lists are rebuilt on every step
    (128, 16384, 8, True, 5, "256"),
    (32, 301, 5, True, 6, None),        # ragged pool, more ranks than some levels have tiles
])
def test_dcgrid_sharded_equals_single_gpu(gpu, d, M, world, solids, steps, unit):
    p = scene_params(d, solids=solids)
    one = FluidSimulationDCGrid((d, d, d), M, p)
    sh = FluidSimulationDCGridSharded((d, d, d), M, p, world, options={"shard_unit": int(unit)} if unit else None)
    _same_state(one, sh, ("density", "velocity", "fluidity"), "after reset")
    for s in range(steps):
        for sim in (one, sh):
            sim.advectVelocity(); sim.adaptTopology(); sim.project()
        if s == steps - 1:
            _same_state(one, sh, ("pressure", "t_pressure", "divergence", "velocity"), f"after project, step {s}")
            assert sh.debugStats() == one.debugStats()
        for sim in (one, sh):
            sim.advectDensity()
    _same_state(one, sh, ("density", "velocity", "fluidity"), "at the end")
    assert abs(sh.totalDensity() - one.totalDensity()) <= 1e-9 * max(1.0, abs(one.totalDensity()))
    assert np.abs(one.field("density")).max() > 0


@pytest.mark.parametrize("size,M,world", [((32, 32, 128), 1500, 4), ((32, 96, 32), 1200, 3), ((128, 32, 32), 1500, 2)])
def test_dcgrid_sharded_slabs_follow_the_longest_axis(gpu, size, M, world):
    """Non-cubic domains: the ranks' slabs are stacked along the longest axis; for y / z stacking the ordered levels
    are renumbered too (dense maps in the mirror).  Results must not depend on any of it."""
    p = scene_params(*size, solids=size[0] <= 32)  # (the sphere's radius scales with gx: it would fill a 128 x 32 x 32 domain)
    one = FluidSimulationDCGrid(size, M, p)
    sh = FluidSimulationDCGridSharded(size, M, p, world, options={"resort_every": 2})
    one.step(9)
    sh.step(9)
    _same_state(one, sh, ("density", "velocity", "fluidity"), f"{size} world {world}")
    assert sh.info("resorts") > 0 and np.abs(one.field("density")).max() > 0
    pos = np.random.default_rng(5).uniform(0, 1, size=(500, 3)).astype(np.float32) * np.float32(size)
    np.testing.assert_array_equal(_bits(one.sampleField("velocity", pos, True)), _bits(sh.sampleField("velocity", pos, True)))


def test_dcgrid_sharded_vs_oracle_graph_path(gpu):
    d, M = 64, 2000
    p = scene_params(d, solids=True)
    sh = FluidSimulationDCGridSharded((d, d, d), M, p, 4, options={"resort_every": 2})  # slab field order re-sorted every other topology change
    orc = Oracle(p, M)
    sh.step(9); orc.step(9)  # reaches the fixed point: the last steps replay the captured graph
    assert sh.counters()[7] == 1
    act = orc.topology()["level"] != 0xFF
    cells = np.repeat(act, 64)
    for f in ("density", "velocity"):
        np.testing.assert_array_equal(_bits(sh.field(f)[cells]), _bits(orc.field(f)[cells]), err_msg=f)
    sh.projectLocal(); orc.project_local()
    np.testing.assert_array_equal(_bits(sh.field("pressure")[cells]), _bits(orc.field("pressure")[cells]))


def test_dcgrid_sharded_rejects_bad_decompositions(gpu):
    p = scene_params(64)
    with pytest.raises(DcgError):
        FluidSimulationDCGridSharded((64, 64, 64), 2000, p, 9)            # more than 8 ranks
    with pytest.raises(DcgError):
        FluidSimulationDCGridSharded((64, 64, 64), 2000, p, 4, rank=1, nlocal=2)   # neither 1 nor world ranks per process
