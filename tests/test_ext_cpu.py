"""CPU properties of the extension SPECIFICATION (oracle/dcgrid_oracle.cpp "EXTENSIONS").  The
reference snapshot has none of these features — parity against it is unpinned — so what can be checked without a GPU
is that the specification is sane: switched off it is the reference, switched on it has the properties the schemes are
defined by.  The CUDA kernels are compared with it bit for bit in tests/test_ext_gpu.py."""
import numpy as np

from dcgrid_b200 import make_ext, scene_params
from tests._oracle import Oracle

D, M = 32, 300  # level 0 is sparse (227 of 512 blocks), levels 1-3 are fully allocated


def run(kw, steps, solids=True, d=D, m=M):
    p = scene_params(d, solids=solids)
    o = Oracle(p, m)
    o.set_ext(make_ext(**kw))
    o.reset()
    o.step(steps)
    return o


def active_cells(o):
    return np.repeat(o.topology(with_apron=False)["level"] != 0xFF, 64)


def leaf_cells(o):
    t = o.topology(with_apron=False)
    return np.repeat(t["level"] != 0xFF, 64) & (np.repeat(t["child"].reshape(-1), 8) == np.uint64(0xFFFFFFFFFFFFFFFF))


def test_all_switches_off_is_the_reference_path():
    a, b = run({}, 5), Oracle(scene_params(D, solids=True), M)
    b.step(5)
    for f in ("density", "velocity", "fluidity"):
        assert np.array_equal(a.field(f).view(np.uint32), b.field(f).view(np.uint32)), f
    assert np.array_equal(a.topology()["apron"], b.topology()["apron"])


def test_maccormack_is_bounded_and_less_diffusive():
    sl, mc = run({}, 10, solids=False), run(dict(advection=1), 10, solids=False)
    rate = sl.params.density_emission_rate
    for o in (sl, mc):
        q = o.field("density")[leaf_cells(o)]
        assert q.min() >= 0.0 and q.max() <= np.float32(rate) * (1 + 1e-6), "clamped to the corners of the forward sample: no new extrema"
    assert not np.array_equal(sl.field("density"), mc.field("density"))
    # less numerical diffusion: the plume keeps more cells close to the emission value
    near = lambda o: int(np.count_nonzero(o.field("density")[leaf_cells(o)] > 0.9 * rate))
    assert near(mc) >= near(sl)


def test_source_pass_conserves_water_and_heats_where_it_condenses():
    o = run(dict(sources=1), 6)
    leaf = leaf_cells(o)
    before = {f: o.field(f).copy() for f in ("density", "vapor", "temperature", "velocity")}
    o.apply_sources()
    after = {f: o.field(f) for f in before}
    water0 = (before["density"] + before["vapor"])[leaf].astype(np.float64)
    water1 = (after["density"] + after["vapor"])[leaf].astype(np.float64)
    np.testing.assert_allclose(water1, water0, rtol=1e-6, atol=1e-9)
    dq = (after["density"] - before["density"])[leaf]
    dth = (after["temperature"] - before["temperature"])[leaf]
    assert dq.max() > 0, "the moist inlet air condenses"
    big = np.abs(dq) > 1e-6  # (the temperature is ~300: increments below its ulp are rounded away)
    assert np.all(np.sign(dth[big]) == np.sign(dq[big])), "latent heat has the sign of the phase change"
    assert after["density"][leaf].min() >= 0.0 and after["vapor"][leaf].min() >= 0.0
    assert np.isfinite(after["velocity"]).all()


def test_flow_driven_score_refines_where_the_flow_is():
    o = run(dict(score_mode=1), 0)
    assert o.level_table()["loads"][0] == 0, "no vorticity, no refinement: only the ordered levels exist after reset"
    o.step(8)
    loads = o.level_table()["loads"]
    assert loads[0] > 0
    t = o.topology(with_apron=False)
    fine = t["pos"][t["level"] == 0]
    # the plume rises from the inlet disc in the middle of the floor: the finest blocks sit around the axis
    c = fine.mean(axis=0)
    assert abs(c[0] - D / 2) < D / 4 and abs(c[2] - D / 2) < D / 4
    w = o.field("vorticity")
    assert np.isfinite(w).all() and np.abs(w[active_cells(o)]).max() > 0


def test_terrain_fluidity():
    kw = dict(terrain=1, terrain_height=10.0, terrain_wavelength=16.0)
    o = run(kw, 0)
    pos = np.stack(np.meshgrid(np.arange(D) + 0.5, [0.5, 9.5, 20.5], np.arange(D) + 0.5, indexing="ij"), axis=-1).reshape(-1, 3).astype(np.float32)
    f = o.sample_field("fluidity", pos).reshape(D, 3, D)
    assert f[:, 2].min() == 1.0, "above the peaks everything is fluid"
    assert f[8, 0, 8] == 0.0 and f[8, 1, 8] < 1.0, "a hill top (x = z = wavelength / 2) is solid up to its height"
    assert f[16, 0, 16] > 0.5 > f[8, 0, 16], "valleys cross at x, z = k * wavelength: the inlet in the domain centre stays open"
    o.step(4)
    assert o.field("density").max() > 0


def test_precise_sampling_reproduces_cell_values_at_cell_centres():
    o = run(dict(sources=1), 5)
    t = o.topology(with_apron=False)
    rng = np.random.default_rng(3)
    slots = rng.choice(np.flatnonzero((t["level"] != 0xFF) & (t["child"] == np.uint64(0xFFFFFFFFFFFFFFFF)).all(axis=1)), 64)
    bits = np.arange(64)
    cx = ((bits >> 5) & 1) * 2 + ((bits >> 2) & 1)
    cy = ((bits >> 4) & 1) * 2 + ((bits >> 1) & 1)
    cz = ((bits >> 3) & 1) * 2 + (bits & 1)
    for b in slots:
        s = float(1 << int(t["level"][b]))
        centres = (t["pos"][b][None, :] + np.stack([cx, cy, cz], axis=1) + 0.5) * s
        for name in ("temperature", "density"):
            want = o.field(name)[b * 64:(b + 1) * 64]
            np.testing.assert_array_equal(o.sample_field(name, centres, precise=False), want)
            np.testing.assert_array_equal(o.sample_field(name, centres, precise=True), want)
