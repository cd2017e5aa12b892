"""Parity of the CUDA uniform-grid solver (through the C ABI) against the CPU oracle.

Bit-exact bar: both sides evaluate the reference's expressions in the reference's order
without FMA contraction (kernels: -fmad=false, oracle: -ffp-contract=off)."""
import numpy as np
import pytest

from dcgrid_b200 import FluidSimulationUniform, scene_params
from tests._oracle import Oracle

pytestmark = pytest.mark.gpu

FIELDS_AFTER_PROJECT = ("pressure", "t_pressure", "divergence")


def _bits(a):
    return np.ascontiguousarray(a, dtype=np.float32).view(np.uint32)


def _run_pair(d, solids, steps, schedule, size=None):
    p = scene_params(d, solids=solids) if size is None else scene_params(*size, solids=solids)
    size = (p.gx, p.gy, p.gz)
    sim = FluidSimulationUniform(size, p)
    orc = Oracle(p)
    if schedule.startswith("jacobi"):
        sim.setJacobiSchedule(2, 1, int(schedule[6:]))
        orc.set_jacobi_schedule(2, 1, int(schedule[6:]))
    for s in range(steps):
        sim.advectVelocity(); orc.advect_velocity()
        sim.adaptTopology(); orc.adapt_topology()
        if schedule == "project":
            sim.project(); orc.project()
        else:
            sim.projectLocal(); orc.project_local()
        if s == steps - 1:
            for f in FIELDS_AFTER_PROJECT:
                np.testing.assert_array_equal(_bits(sim.field(f)), _bits(orc.field(f)), err_msg=f"{f} after project, step {s}")
        sim.advectDensity(); orc.advect_density()
    return sim, orc


@pytest.mark.parametrize("d,solids,steps,schedule", [
    (32, False, 12, "project"),
    (32, True, 12, "local"),
    (64, False, 25, "project"),
    (64, True, 10, "jacobi25"),
])
def test_uniform_bit_exact_vs_oracle(gpu, d, solids, steps, schedule):
    sim, orc = _run_pair(d, solids, steps, schedule)
    for f in ("density", "velocity", "fluidity"):
        np.testing.assert_array_equal(_bits(sim.field(f)), _bits(orc.field(f)), err_msg=f)
    assert sim.debugStats() == orc.debug_stats()
    assert orc.field("density").max() > 0  # the inlet actually injected smoke


def test_uniform_non_cubic(gpu):
    sim, orc = _run_pair(32, True, 8, "project", size=(32, 64, 48))
    for f in ("density", "velocity"):
        np.testing.assert_array_equal(_bits(sim.field(f)), _bits(orc.field(f)), err_msg=f)


@pytest.mark.parametrize("size,solids,schedule,options", [
    ((256, 16, 32), False, "project", None),    # levels 0 and 1 take the z-marching kernels (x extent in whole warps of float4)
    ((128, 64, 72), True, "local", None),       # z extent not a multiple of the chunk; solids; projectLocal's 5 pairs
    ((256, 16, 32), False, "project", {"stencil": 1}),  # the one-thread-per-cell kernels at the same size
    ((128, 64, 72), True, "project", {"experiment": 64}),       # the one-row divergence kernel (sizes with gy % 16 != 0 take it)
    ((128, 64, 72), True, "project", {"experiment": 1024}),     # two-row divergence without the fused level-1 restriction
    ((128, 64, 72), True, "project", None),                     # ... and with it (the shipped choice on one GPU)
])
def test_uniform_z_marching_kernels(gpu, size, solids, schedule, options):
    """The z-marching kernels only engage from 128 cells in x (two-row divergence: gy % 16 == 0; its fused level-1 restriction:
    project() on one GPU): sizes the other cases never reach."""
    p = scene_params(*size, solids=solids)
    sim = FluidSimulationUniform(size, p, options=options)
    orc = Oracle(p)
    for s in range(6):
        sim.advectVelocity(); orc.advect_velocity()
        sim.adaptTopology(); orc.adapt_topology()
        if schedule == "project":
            sim.project(); orc.project()
        else:
            sim.projectLocal(); orc.project_local()
        for f in FIELDS_AFTER_PROJECT + ("velocity",):
            np.testing.assert_array_equal(_bits(sim.field(f)), _bits(orc.field(f)), err_msg=f"{f} after project, step {s}")
        sim.advectDensity(); orc.advect_density()
    for f in ("density", "velocity"):
        np.testing.assert_array_equal(_bits(sim.field(f)), _bits(orc.field(f)), err_msg=f)
    assert orc.field("density").max() > 0


def test_uniform_step_graph_equals_calls(gpu):
    """dcg_step (CUDA-graph replay of the 4-call sequence) == the four calls issued one by one."""
    p = scene_params(32)
    a = FluidSimulationUniform((32, 32, 32), p)
    b = FluidSimulationUniform((32, 32, 32), p)
    a.step(7)
    for _ in range(7):
        b.advectVelocity(); b.adaptTopology(); b.project(); b.advectDensity()
    for f in ("density", "velocity"):
        np.testing.assert_array_equal(_bits(a.field(f)), _bits(b.field(f)))
    assert a.lastStepMs() > 0


def test_uniform_reset_restores_initial_state(gpu):
    p = scene_params(32)
    a = FluidSimulationUniform((32, 32, 32), p)
    a.step(5)
    assert np.abs(a.field("velocity")).max() > 0
    a.reset()
    assert np.abs(a.field("velocity")).max() == 0 and np.abs(a.field("density")).max() == 0
    a.step(5)
    b = FluidSimulationUniform((32, 32, 32), p)
    b.step(5)
    np.testing.assert_array_equal(_bits(a.field("density")), _bits(b.field("density")))


@pytest.mark.parametrize("no_fuse", [0, 1])
def test_uniform_speculative_velocity_any_call_order(gpu, no_fuse):
    """advectDensity() also writes the next step's advected velocity into the idle ping-pong buffer
    (k_u_advect_both); advectVelocity() may only use it if nothing touched the state in between."""
    size = (32, 32, 32)
    p = scene_params(32, solids=True)
    sim = FluidSimulationUniform(size, p, options={"advect_no_fuse": no_fuse})
    orc = Oracle(p)
    sim.step(3); orc.step(3)
    seq = ["advect_density", "project", "advect_velocity", "advect_density", "advect_velocity", "advect_velocity",
           "advect_density", "project_local", "advect_density", "advect_velocity", "advect_density", "set_params",
           "advect_velocity", "project", "advect_density"]
    names = {"advect_density": "advectDensity", "advect_velocity": "advectVelocity", "project": "project", "project_local": "projectLocal"}
    for i, op in enumerate(seq):
        if op == "set_params":
            p.dt = 2.5
            sim.setParams(p); orc.set_params(p)
        else:
            getattr(sim, names[op])(); getattr(orc, op)()
        for f in ("density", "velocity"):
            np.testing.assert_array_equal(np.ascontiguousarray(sim.field(f)).view(np.uint32), np.ascontiguousarray(orc.field(f)).view(np.uint32),
                                          err_msg=f"{f} after call {i} ({op})")
    sim.step(4); orc.step(4)
    for f in ("density", "velocity"):
        np.testing.assert_array_equal(np.ascontiguousarray(sim.field(f)).view(np.uint32), np.ascontiguousarray(orc.field(f)).view(np.uint32), err_msg=f)
