"""z-slab decomposition of the uniform solver: an N-rank run must equal the 1-GPU run bit for bit (same
per-cell arithmetic, only the owner of the memory differs).  Here all ranks live in one process on one
device (nlocal == world), which exercises every ownership / peer-pointer index path of
uniform.cu (slab decomposition); the multi-process NVLink path is tests/mgpu_uniform_check.py (torchrun, >= 2 GPUs)."""
import numpy as np
import pytest

from dcgrid_b200 import DcgError, FluidSimulationUniform, FluidSimulationUniformSharded, scene_params
from tests._oracle import Oracle

pytestmark = pytest.mark.gpu


def _bits(a):
    return np.ascontiguousarray(a, dtype=np.float32).view(np.uint32)


@pytest.mark.parametrize("size,world,solids,steps,schedule", [
    ((32, 32, 32), 2, False, 8, "project"),
    ((32, 32, 32), 4, True, 8, "project"),
    ((32, 32, 32), 8, True, 6, "local"),      # slab of 4 planes: coarse levels live on a subset of the ranks
    ((64, 32, 48), 3, True, 6, "project"),    # slab 16, non-cubic, world not a power of two
    ((64, 64, 64), 8, False, 10, "project"),
])
def test_sharded_equals_single_gpu_and_oracle(gpu, size, world, solids, steps, schedule):
    p = scene_params(*size, solids=solids)
    one = FluidSimulationUniform(size, p)
    sh = FluidSimulationUniformSharded(size, p, world)
    orc = Oracle(p)
    assert sh.numCells == one.numCells
    for s in range(steps):
        for sim in (one, sh):
            sim.advectVelocity(); sim.adaptTopology()
            sim.project() if schedule == "project" else sim.projectLocal()
        orc.advect_velocity(); orc.adapt_topology()
        orc.project() if schedule == "project" else orc.project_local()
        if s == steps - 1:
            for f in ("pressure", "t_pressure", "divergence"):
                np.testing.assert_array_equal(_bits(sh.field(f)), _bits(one.field(f)), err_msg=f"{f} (vs 1 GPU)")
                np.testing.assert_array_equal(_bits(sh.field(f)), _bits(orc.field(f)), err_msg=f"{f} (vs oracle)")
        for sim in (one, sh):
            sim.advectDensity()
        orc.advect_density()
    for f in ("density", "velocity", "fluidity"):
        np.testing.assert_array_equal(_bits(sh.field(f)), _bits(one.field(f)), err_msg=f)
        np.testing.assert_array_equal(_bits(sh.field(f)), _bits(orc.field(f)), err_msg=f"{f} (vs oracle)")
    assert sh.debugStats() == one.debugStats() == orc.debug_stats()
    assert abs(sh.totalDensity() - one.totalDensity()) <= 1e-9 * max(1.0, abs(one.totalDensity()))
    assert orc.field("density").max() > 0


def test_sharded_step_and_reset(gpu):
    size = (32, 32, 32)
    p = scene_params(32, solids=True)
    a = FluidSimulationUniformSharded(size, p, 4)
    b = FluidSimulationUniform(size, p)
    a.step(5); b.step(5)
    np.testing.assert_array_equal(_bits(a.field("density")), _bits(b.field("density")))
    a.reset(); b.reset()
    a.step(3); b.step(3)
    np.testing.assert_array_equal(_bits(a.field("velocity")), _bits(b.field("velocity")))
    assert a.lastStepMs() > 0


def test_sharded_rejects_bad_decompositions(gpu):
    p = scene_params(32)
    with pytest.raises(DcgError):
        FluidSimulationUniformSharded((32, 32, 32), p, 5)      # 32 % 5 != 0
    with pytest.raises(DcgError):
        FluidSimulationUniformSharded((32, 32, 32), p, 16)     # more than 8 ranks
