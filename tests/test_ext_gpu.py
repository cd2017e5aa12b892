"""The extensions beyond the reference snapshot (flow-driven score, temperature / vapor with the fused
source pass, MacCormack, terrain SDF, point sampling, state dump / load) — CUDA through the C ABI against their CPU
specification in oracle/dcgrid_oracle.cpp.  PARITY UNPINNED against the reference: the snapshot under /root/reference
contains none of these features, the oracle is the only pin.  The bar is the same as everywhere else:
bit-exact block pool and bit-exact fp32 fields on every cell of every active block."""
import numpy as np
import pytest

from dcgrid_b200 import DcgError, FluidSimulationDCGrid, FluidSimulationUniform, make_ext, scene_params
from tests._oracle import Oracle
from tests.test_dcgrid_gpu import _bits, assert_same_fields, assert_same_topology

pytestmark = pytest.mark.gpu

TERRAIN = dict(terrain=1, terrain_height=12.0, terrain_wavelength=16.0)
CASES = {
    "sources": dict(sources=1),
    "maccormack": dict(advection=1),
    "flow_score": dict(score_mode=1),
    "terrain": dict(TERRAIN),
    "sources_maccormack": dict(sources=1, advection=1),
    "cloud_scene": dict(sources=1, advection=1, score_mode=1, **TERRAIN),  # BASELINE configs[2]/[3] in small
    # total-order selection evaluated entirely on the device (radix sorts + ballot / prefix-sum compaction)
    "device_selection_geometric": dict(selection=1),
    "device_selection_flow": dict(selection=1, score_mode=1),
    "wildfire_scene": dict(selection=1, score_mode=1, sources=1, advection=1),  # BASELINE configs[3] in small
}


def fields_of(kw):
    names = ["density", "velocity", "fluidity"]
    if kw.get("sources"):
        names += ["temperature", "vapor"]
    if kw.get("sources") or kw.get("score_mode"):
        names += ["vorticity"]
    return names


def make_pair(d, M, kw, options=None):
    p = scene_params(d, solids=True)
    e = make_ext(**kw)
    sim = FluidSimulationDCGrid((d, d, d), M, p, options=options)
    sim.setExt(e)
    sim.reset()
    orc = Oracle(p, M)
    orc.set_ext(e)
    orc.reset()
    return sim, orc


@pytest.mark.parametrize("case", sorted(CASES))
def test_extension_matches_its_cpu_specification(gpu, case):
    kw = CASES[case]
    sim, orc = make_pair(64, 4096, kw)
    host0 = sim.info("host_selections")  # (the constructor's own reset ran before the extension was switched on)
    act = assert_same_topology(sim, orc, "after reset")
    assert_same_fields(sim, orc, [f for f in fields_of(kw) if f != "vorticity"], act, "after reset")
    for s in range(8):
        sim.advectVelocity(); orc.advect_velocity()
        sim.adaptTopology(); orc.adapt_topology()
        if kw.get("sources"):
            sim.applySources(); orc.apply_sources()
        sim.project(); orc.project()
        act = assert_same_topology(sim, orc, f"step {s}")
        assert_same_fields(sim, orc, ("pressure", "t_pressure", "divergence", "velocity"), act, f"after project, step {s}")
        sim.advectDensity(); orc.advect_density()
        assert_same_fields(sim, orc, fields_of(kw), act, f"step {s}")
    if kw.get("score_mode"):
        assert orc.level_table()["loads"][0] > 0 and sim.counters()[7] == 0, "flow-driven refinement follows the plume and never reaches a fixed point"
    if kw.get("selection"):
        assert sim.info("device_selections") > 0 and sim.info("host_selections") == host0 and sim.counters()[2] > 0, "blocks moved, no host selection"
    if kw.get("sources"):
        assert orc.field("density").max() > 0 and float(np.ptp(orc.field("temperature")[np.repeat(act, 64)])) > 1.0


@pytest.mark.parametrize("options", [{"no_resort": 1}, {"resort_every": 1}, {"advect": 1, "stencil": 1, "jacobi": 1}, {"jacobi": 2, "jacobi_max_ctas": 8}])
def test_extensions_under_kernel_variants_and_dcg_step(gpu, options):
    """dcg_step (the call a user makes; inserts the source pass itself) under the layout / kernel variants."""
    kw = CASES["cloud_scene"]
    sim, orc = make_pair(64, 4096, kw, options=options)
    sim.step(7)
    orc.step(7)
    act = assert_same_topology(sim, orc, str(options))
    assert_same_fields(sim, orc, fields_of(kw), act, str(options))


def test_sources_with_the_geometric_score_reach_the_fixed_point_and_replay_the_graph(gpu):
    kw = dict(sources=1)
    sim, orc = make_pair(32, 300, kw)
    n = 0
    while sim.counters()[7] == 0 and n < 400:
        sim.step(10)
        n += 10
    sim.step(6)
    orc.step(n + 6)
    act = assert_same_topology(sim, orc)
    assert_same_fields(sim, orc, fields_of(kw), act)
    c = sim.counters()
    assert c[7] == 1 and c[4] > 0, f"geometric score + sources: the topology settles ({n} steps) and the step graph (incl. the source pass) is replayed"
    assert sim.info("graph_launches_per_step") > 0


def test_point_sampling_matches_samplecoarse_and_sampleprecise(gpu):
    kw = dict(sources=1)
    sim, orc = make_pair(64, 4096, kw)
    sim.step(6)
    orc.step(6)
    rng = np.random.default_rng(7)
    pos = rng.uniform(0.0, 64.0, size=(4000, 3)).astype(np.float32)
    for name in ("density", "velocity", "temperature", "fluidity"):
        for precise in (False, True):
            np.testing.assert_array_equal(_bits(sim.sampleField(name, pos, precise)), _bits(orc.sample_field(name, pos, precise)), err_msg=f"{name} precise={precise}")
    # the dense level-0 resampling is sampleCoarse at the level-0 cell centres
    zz, yy, xx = np.meshgrid(np.arange(64), np.arange(64), np.arange(64), indexing="ij")
    centres = np.stack([xx, yy, zz], axis=-1).reshape(-1, 3).astype(np.float32) + 0.5
    np.testing.assert_array_equal(_bits(sim.sampleField("density", centres)), _bits(sim.field("density", layout=1, count=64 ** 3)))


@pytest.mark.parametrize("kw", [{}, CASES["cloud_scene"]], ids=["reference_features", "cloud_scene"])
def test_state_dump_and_load_continue_bit_identically(gpu, tmp_path, kw):
    d, M = 64, 4096
    p = scene_params(d, solids=True)
    a = FluidSimulationDCGrid((d, d, d), M, p)
    if kw:
        a.setExt(make_ext(**kw))
        a.reset()
    a.step(9)
    path = tmp_path / "state.dcg"
    a.saveState(path)
    a.step(5)
    b = FluidSimulationDCGrid((d, d, d), M, p, options={"resort_every": 3})  # a different layout history: results must not depend on it
    b.loadState(path)
    assert bytes(b.getExt()) == bytes(a.getExt())
    b.step(5)
    ta, tb = a.topology(), b.topology()
    for k in ("level", "pos", "parent", "child", "apron"):
        act = ta["level"] != 0xFF
        np.testing.assert_array_equal(ta[k][act], tb[k][act], err_msg=k)
    cells = np.repeat(act, 64)
    for f in fields_of(kw):
        np.testing.assert_array_equal(_bits(a.field(f)[cells]), _bits(b.field(f)[cells]), err_msg=f)
    np.testing.assert_array_equal(a.counters()[:4], b.counters()[:4])
    with pytest.raises(DcgError):
        FluidSimulationDCGrid((32, 32, 32), 300, scene_params(32)).loadState(path)  # another grid / pool size


def test_extension_errors(gpu):
    u = FluidSimulationUniform((16, 16, 16), scene_params(16))
    with pytest.raises(DcgError):
        u.setExt(make_ext(sources=1))
    u.setExt(make_ext())  # every switch off: accepted everywhere
    s = FluidSimulationDCGrid((32, 32, 32), 300, scene_params(32))
    with pytest.raises(DcgError):
        s.field("temperature")  # not switched on
    with pytest.raises(DcgError):
        s.setExt(make_ext(terrain=1, terrain_wavelength=0.0))
    s.step(2)  # the failed call left the instance usable and unchanged
    o = Oracle(scene_params(32), 300)
    o.step(2)
    act = assert_same_topology(s, o)
    assert_same_fields(s, o, ("density", "velocity"), act)
