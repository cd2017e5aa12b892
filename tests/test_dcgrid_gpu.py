"""Parity of the CUDA DCGrid solver (through the C ABI) against the CPU oracle.

The oracle is pinned bit-for-bit against the reference's own CUDA kernels (tests/golden/summary.json),
so equality with the oracle is equality with the reference under rank-order slot allocation.
Integer structures (block pool, parent/child links, the 6^3 apron map, level loads) are compared
raw, slot by slot; floating-point fields are compared bit for bit."""
import numpy as np
import pytest

from dcgrid_b200 import DcgError, FluidSimulationDCGrid, scene_params
from tests._oracle import Oracle

pytestmark = pytest.mark.gpu


def _bits(a):
    return np.ascontiguousarray(a, dtype=np.float32).view(np.uint32)


def assert_same_topology(sim, orc, where=""):
    st, ot = sim.topology(), orc.topology()
    np.testing.assert_array_equal(st["level"], ot["level"], err_msg=f"levels {where}")
    act = ot["level"] != 0xFF
    np.testing.assert_array_equal(st["pos"][act], ot["pos"][act], err_msg=f"positions {where}")
    np.testing.assert_array_equal(st["parent"][act], ot["parent"][act], err_msg=f"parent {where}")
    np.testing.assert_array_equal(st["child"][act], ot["child"][act], err_msg=f"children {where}")
    np.testing.assert_array_equal(st["apron"][act], ot["apron"][act], err_msg=f"apron {where}")
    np.testing.assert_array_equal(sim.levelTable()["loads"], orc.level_table()["loads"], err_msg=f"loads {where}")
    return act


def assert_same_fields(sim, orc, names, act, where=""):
    cells = np.repeat(act, 64)
    for f in names:
        a, b = sim.field(f), orc.field(f)
        if f == "velocity":
            a, b = a[cells], b[cells]
        else:
            a, b = a[cells], b[cells]
        np.testing.assert_array_equal(_bits(a), _bits(b), err_msg=f"{f} {where}")


def run_pair(d, M, solids, steps, schedule="project", check_every=1, options=None):
    p = scene_params(d, solids=solids)
    sim = FluidSimulationDCGrid((d, d, d), M, p, options=options)
    orc = Oracle(p, M)
    assert sim.levels == orc.levels and sim.sparseLevels == orc.sparse_levels
    act = assert_same_topology(sim, orc, "after reset")
    assert_same_fields(sim, orc, ("density", "velocity", "fluidity"), act, "after reset")
    for s in range(steps):
        sim.advectVelocity(); orc.advect_velocity()
        sim.adaptTopology(); orc.adapt_topology()
        if schedule == "project":
            sim.project(); orc.project()
        else:
            sim.projectLocal(); orc.project_local()
        if s % check_every == 0 or s == steps - 1:
            act = assert_same_topology(sim, orc, f"step {s}")
            assert_same_fields(sim, orc, ("pressure", "t_pressure", "divergence", "velocity"), act, f"after project, step {s}")
        if s == steps - 1:
            # the residual is only defined between project() and advectDensity(): in the reference
            # `divergence` aliases `t_density` (dcgrid_structure.cu:94-102)
            assert sim.debugStats() == orc.debug_stats()
        sim.advectDensity(); orc.advect_density()
        if s % check_every == 0 or s == steps - 1:
            assert_same_fields(sim, orc, ("density", "velocity", "fluidity"), act, f"end of step {s}")
    return sim, orc


@pytest.mark.parametrize("d,M,solids,steps,schedule", [
    (32, 300, False, 12, "project"),
    (32, 300, True, 8, "local"),
    (64, 4096, False, 14, "project"),   # perpetual 2-block move cycle: adaptation runs every step
    (64, 2000, True, 12, "project"),
    (128, 16384, True, 6, "project"),
])
def test_dcgrid_bit_exact_vs_oracle(gpu, d, M, solids, steps, schedule):
    sim, orc = run_pair(d, M, solids, steps, schedule)
    c, oc = sim.counters(), orc.counters()
    assert c[2] == oc[2] and c[3] == oc[3], "moved / refined counts"
    assert c[5] == 0, "failed allocations"
    assert orc.field("density").max() > 0


@pytest.mark.parametrize("options", [
    {"jacobi": 2},                                         # TMA-ring Jacobi on every level, incl. ragged last tiles
    {"jacobi": 2, "no_snake": 1},
    {"jacobi": 4},                                         # the 4-cells-per-thread ring kernel
    {"resort_every": 1, "jacobi": 2},                      # field order re-sorted by position after EVERY topology change
    {"resort_every": 3},                                   # incremental mirror updates between the re-sorts
    {"no_resort": 1},                                      # field order = the reference's slot order throughout
    {"jacobi": 1, "advect": 1},                            # one-CTA-per-tile kernels
    {"jacobi_ctas_per_sm": 1, "advect_ctas_per_sm": 1},    # one resident CTA per SM: every CTA walks several tiles
    {"advect_no_fuse": 1, "advect_slot_order": 1},         # density and velocity advection as separate passes, pool order
    {"advect_min_blocks": 3, "stencil": 1},                # 3-CTA/SM advection build; one-CTA-per-tile divergence / gradient
    {"no_pdl": 1, "host_selection": 1},                    # plain stream order; the reference's host selection on every level
    {"coarse_in_gmem": 1, "zero_all": 1, "apply_min_blocks": 3, "advect_min_blocks": 4},
    {"experiment": 32},                                    # adaptTopology computes its scores on entry (no early scores on the side stream)
    {"jacobi": 2, "experiment": 9},                        # the sharded defaults on one GPU: restriction as ONE walk up the block tree
                                                           # (k_dc_restrict_tree), prolongation by parent block (k_dc_prolongate_parents)
], ids=lambda o: ",".join(f"{k}={v}" for k, v in o.items()))
@pytest.mark.parametrize("d,M,solids,steps", [(64, 2000, True, 8), (32, 301, False, 6), (128, 16384, True, 4)])
def test_dcgrid_kernel_variants_bit_exact(gpu, options, d, M, solids, steps):
    """The persistent cp.async.bulk (TMA ring) kernels and the one-CTA-per-tile kernels are interchangeable:
    each combination must reproduce the oracle bit for bit (variants are creation-time options of the instance,
    struct dcg_options in include/dcgrid_b200.h)."""
    run_pair(d, M, solids, steps, "project", check_every=steps, options=options)


@pytest.mark.parametrize("size,M,steps", [((64, 64, 128), 4000, 10), ((128, 64, 64), 3000, 8), ((32, 64, 96), 1500, 8)])
def test_dcgrid_non_cubic_grids_bit_exact_vs_oracle(gpu, size, M, steps):
    """Non-cubic domains (the multi-GPU bench runs one 512 x 512 x 512N scene): level count from the smallest
    dimension, level maps and ordered-level indexing with three different extents."""
    p = scene_params(*size, solids=True)
    sim = FluidSimulationDCGrid(size, M, p)
    orc = Oracle(p, M)
    assert sim.levels == orc.levels and sim.sparseLevels == orc.sparse_levels
    sim.step(steps); orc.step(steps)
    act = assert_same_topology(sim, orc, f"{size} after {steps} steps")
    assert_same_fields(sim, orc, ("density", "velocity", "fluidity"), act, f"{size}")
    sim.project(); orc.project()
    assert_same_fields(sim, orc, ("pressure", "divergence"), act, f"{size} after project")
    assert orc.field("density").max() > 0


def test_dcgrid_steady_state_skip_and_graph(gpu):
    """Once the (topology, moveLimit) fixed point is proven adaptTopology is skipped and dcg_step replays a
    CUDA graph; results must equal the call-by-call path and the oracle."""
    d, M = 64, 2000
    p = scene_params(d, solids=True)
    a = FluidSimulationDCGrid((d, d, d), M, p)
    orc = Oracle(p, M)
    a.step(9)
    orc.step(9)
    c = a.counters()
    assert c[7] == 1 and c[4] > 0, "steady state reached and adaptTopology skipped"
    act = assert_same_topology(a, orc, "after 9 steps")
    assert_same_fields(a, orc, ("density", "velocity"), act, "graph path")
    assert a.lastStepMs() > 0


def test_dcgrid_speculative_velocity_is_dropped_when_state_changes(gpu):
    """advectDensity() also writes the next step's advected velocity into the idle ping-pong buffer
    (k_dc_advect_pipe<2>); advectVelocity() may only use it if nothing touched the state in between.  Drive the
    nine virtuals in orders the reference's UI never uses and compare with the oracle after every call."""
    d, M = 64, 2000
    p = scene_params(d, solids=True)
    sim = FluidSimulationDCGrid((d, d, d), M, p)
    orc = Oracle(p, M)
    sim.step(3); orc.step(3)
    seq = ["advect_density", "project", "advect_velocity",          # project between producer and consumer
           "advect_density", "advect_velocity", "advect_velocity",  # consumed once, second call runs the kernel
           "advect_density", "adapt_topology", "advect_velocity",
           "advect_density", "project_local", "advect_density", "advect_velocity",
           "advect_density", "set_params", "advect_velocity", "project", "advect_density"]
    names = {"advect_density": "advectDensity", "advect_velocity": "advectVelocity", "project": "project",
             "project_local": "projectLocal", "adapt_topology": "adaptTopology"}
    for i, op in enumerate(seq):
        if op == "set_params":
            p.dt = 2.5
            sim.setParams(p); orc.set_params(p)
        else:
            getattr(sim, names[op])(); getattr(orc, op)()
        act = assert_same_topology(sim, orc, f"call {i} ({op})")
        assert_same_fields(sim, orc, ("density", "velocity"), act, f"call {i} ({op})")
    sim.step(4); orc.step(4)  # back on the graph / fused path
    act = assert_same_topology(sim, orc, "after the sequence")
    assert_same_fields(sim, orc, ("density", "velocity"), act, "after the sequence")


def test_dcgrid_lookup_and_dense_resample(gpu):
    d, M = 64, 2000
    p = scene_params(d, solids=True)
    sim = FluidSimulationDCGrid((d, d, d), M, p)
    orc = Oracle(p, M)
    sim.step(3); orc.step(3)
    rng = np.random.default_rng(0)
    pos = rng.integers(0, d, size=(4096, 3)).astype(np.int32)
    s1, l1 = sim.lookupBlocks(pos)
    s2, l2 = orc.lookup_blocks(pos)
    np.testing.assert_array_equal(s1, s2)
    np.testing.assert_array_equal(l1, l2)
    dense = sim.denseField("density").reshape(d, d, d)  # [z][y][x]
    q = orc.field("density")
    topo = orc.topology(with_apron=False)
    for (x, y, z), slot, lvl in list(zip(pos, s2, l2))[:512]:
        bp = topo["pos"][slot]
        cx, cy, cz = (x >> lvl) - bp[0], (y >> lvl) - bp[1], (z >> lvl) - bp[2]
        bits = ((cx >> 1) << 5) | ((cy >> 1) << 4) | ((cz >> 1) << 3) | ((cx & 1) << 2) | ((cy & 1) << 1) | (cz & 1)
        assert dense[z, y, x] == q[int(slot) * 64 + bits]


def test_dcgrid_pool_too_small_is_an_error_not_an_exit(gpu):
    p = scene_params(256)
    with pytest.raises(DcgError):
        FluidSimulationDCGrid((256, 256, 256), 0, p)
    # 64^3 has 5 levels; the coarsest takes the only block, nothing is left for level 0
    # ("Too few blocks to reach highest resolution", fluid_simulation_dcgrid.cu:50-53)
    with pytest.raises(DcgError):
        FluidSimulationDCGrid((64, 64, 64), 1, scene_params(64))


def test_dcgrid_reset_is_reproducible(gpu):
    d, M = 64, 4096
    p = scene_params(d)
    a = FluidSimulationDCGrid((d, d, d), M, p)
    t0 = a.topology()
    a.step(5)
    a.reset()
    t1 = a.topology()
    for k in ("pos", "level", "parent", "child", "apron"):
        np.testing.assert_array_equal(t0[k], t1[k], err_msg=k)
    assert np.abs(a.field("velocity")).max() == 0


def test_rejected_set_params_leaves_the_instance_intact(gpu):
    """A grid-size change is rejected (the pool is sized at construction, fluid_simulation_dcgrid.cu:9-140) and must
    not leave half-applied parameters behind: the next steps still match the oracle, and repeating the rejected
    call is rejected again (not swallowed by the "identical bytes" shortcut)."""
    d, M = 32, 300
    p = scene_params(d, solids=True)
    sim = FluidSimulationDCGrid((d, d, d), M, p)
    orc = Oracle(p, M)
    sim.step(2); orc.step(2)
    bad = scene_params(64, solids=True)
    for _ in range(2):
        with pytest.raises(DcgError):
            sim.setParams(bad)
    sim.step(3); orc.step(3)
    act = assert_same_topology(sim, orc, "after a rejected set_params")
    assert_same_fields(sim, orc, ("density", "velocity"), act, "after a rejected set_params")
