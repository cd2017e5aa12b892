"""CPU-only checks of the drop-in boundary: the C-ABI library loads, exports every symbol that
include/dcgrid_b200.h declares, keeps SimParams layout-identical to the reference
(src/data/sim_params.h:14-47), and refuses to run without a GPU (no CPU fallback)."""
import ctypes
import os
import re
import subprocess

import pytest

from dcgrid_b200 import _lib
from dcgrid_b200.params import Options, SimParams, default_params, make_options, scene_params

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "dcgrid_b200.h")


def _no_gpu():
    L = _lib.load()
    h = ctypes.c_void_p()
    p = scene_params(8)
    rc = L.dcg_create_uniform(ctypes.byref(p), 0, ctypes.byref(h))
    if rc == 0:
        L.dcg_destroy(h)
    return rc != 0


def declared_symbols():
    text = open(HEADER).read()
    return sorted(set(re.findall(r"DCG_API\s+[\w\s\*]+?\b(dcg_\w+)\s*\(", text)))


def test_header_declares_the_reference_virtuals():
    names = declared_symbols()
    # the nine virtuals of src/fluid_simulation.h:9-22 + lifetime + the additions
    for n in ("dcg_init", "dcg_reset", "dcg_adapt_topology", "dcg_advect_velocity", "dcg_project", "dcg_project_local",
              "dcg_advect_density", "dcg_render", "dcg_debug_stats", "dcg_create_uniform", "dcg_create_dcgrid", "dcg_destroy",
              "dcg_step", "dcg_set_params", "dcg_get_field", "dcg_get_topology", "dcg_last_error"):
        assert n in names, n
    assert len(names) >= 30


def test_library_exports_every_declared_symbol():
    L = _lib.load()
    for n in declared_symbols():
        assert hasattr(L, n), f"libdcgrid_b200.so does not export {n}"
    # and the python binding types every one of them
    assert set(declared_symbols()) == set(_lib.SYMBOLS), set(declared_symbols()) ^ set(_lib.SYMBOLS)


def test_sim_params_layout_matches_reference_struct():
    # float dt; int gx,gy,gz; float dx,rdx, 3 floats; 4 bools; enum; float; float; 4 x float3  == 100 bytes
    assert ctypes.sizeof(SimParams) == 100
    off = {f[0]: getattr(SimParams, f[0]).offset for f in SimParams._fields_}
    assert off["dt"] == 0 and off["gx"] == 4 and off["dx"] == 16 and off["rdx"] == 20
    assert off["velocity_emission_rate"] == 24 and off["emission_radius"] == 32
    assert off["enable_additional_solids"] == 36 and off["render_channel"] == 40


def test_default_params_are_the_reference_defaults():
    L = _lib.load()
    p = SimParams()
    assert L.dcg_default_params(ctypes.byref(p)) == 0
    # src/data/sim_params.cpp:4-34, rdx as src/main.cpp:17
    assert (p.gx, p.gy, p.gz) == (128, 128, 128)
    assert p.dt == 3.0 and p.velocity_emission_rate == 150.0 and p.emission_radius == 750.0
    assert abs(p.dx - 10000.0 / 128) < 1e-4 and abs(p.rdx * p.dx - 1.0) < 1e-6
    assert not p.enable_additional_solids
    q = default_params()
    assert bytes(p) == bytes(q)


def test_options_struct_matches_the_header():
    """struct dcg_options: the ctypes mirror has the size the library reports, a zeroed struct is the default."""
    L = _lib.load()
    o = Options()
    assert L.dcg_default_options(ctypes.byref(o)) == 0
    assert o.struct_size == ctypes.sizeof(Options)
    assert all(getattr(o, f[0]) == 0 for f in Options._fields_ if f[0] not in ("struct_size", "reserved"))
    text = open(HEADER).read()
    body = text[text.index("typedef struct dcg_options {"):text.index("} dcg_options;")]
    declared = re.findall(r"^\s+u?int32_t\s+(\w+)", body, flags=re.M)
    assert declared == [f[0] for f in Options._fields_], "field order of Options vs. struct dcg_options"
    with pytest.raises(KeyError):
        make_options({"no_such_option": 1})
    assert make_options({"jacobi": 2}).jacobi == 2


def test_ext_params_struct_matches_the_header():
    """struct dcg_ext_params: same field order and types as the ctypes mirror; the defaults switch every extension off."""
    from dcgrid_b200.params import ExtParams, make_ext

    e = make_ext()
    assert e.struct_size == ctypes.sizeof(ExtParams)
    assert (e.score_mode, e.advection, e.sources, e.terrain, e.selection) == (0, 0, 0, 0, 0)
    assert e.ambient_temperature > 0 and e.terrain_wavelength > 0
    text = open(HEADER).read()
    body = text[text.index("typedef struct dcg_ext_params {"):text.index("} dcg_ext_params;")]
    declared = re.findall(r"^\s+(?:u?int32_t|float)\s+(\w+)", body, flags=re.M)
    assert declared == [f[0] for f in ExtParams._fields_], "field order of ExtParams vs. struct dcg_ext_params"
    with pytest.raises(KeyError):
        make_ext(no_such_field=1)


def test_null_handles_are_rejected_not_dereferenced():
    L = _lib.load()
    assert L.dcg_step(None, 1) != 0
    assert L.dcg_reset(None) != 0
    assert L.dcg_num_cells(None) == 0
    assert L.dcg_version().startswith(b"dcgrid_b200")


def test_no_cpu_fallback_without_a_gpu():
    if not _no_gpu():
        pytest.skip("a CUDA device is present")
    L = _lib.load()
    h = ctypes.c_void_p()
    p = scene_params(16)
    rc = L.dcg_create_dcgrid(ctypes.byref(p), 300, 0, ctypes.byref(h))
    assert rc == 5  # DCG_ERR_NO_DEVICE
    assert not h
    assert b"no CPU fallback" in L.dcg_last_error(None)
    from dcgrid_b200 import DcgError, FluidSimulationUniform

    with pytest.raises(DcgError):
        FluidSimulationUniform((16, 16, 16), p)


def test_cpp_adapter_fails_loudly_without_a_gpu():
    exe = os.path.join(ROOT, "oracle", "_ref", "adapter_demo")
    if not os.path.exists(exe):
        pytest.skip("oracle/_ref/adapter_demo not built (needs /root/reference at build time)")
    if not _no_gpu():
        pytest.skip("a CUDA device is present")
    r = subprocess.run([exe, "d=16", "M=300"], capture_output=True, text=True, timeout=60)
    assert r.returncode != 0 and "no CPU fallback" in r.stderr
