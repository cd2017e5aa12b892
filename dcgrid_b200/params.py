"""Host-side mirror of the reference's ``struct SimParams`` (src/data/sim_params.h:14-47).

Layout-identical ctypes structure (100 bytes) used on both sides of the C ABI
(``include/dcgrid_b200.h``).  ``default_params`` follows ``SimParams::defaultParams``
(src/data/sim_params.cpp:4-34) and ``scene_params`` the presets of src/data/scenes.h:13-32;
``rdx`` is set here the way src/main.cpp:17 does it (``rdx = 1.f / dx``).
"""
import ctypes

import numpy as np


class Float3(ctypes.Structure):
    _fields_ = [("x", ctypes.c_float), ("y", ctypes.c_float), ("z", ctypes.c_float)]


class SimParams(ctypes.Structure):
    _fields_ = [
        ("dt", ctypes.c_float),
        ("gx", ctypes.c_int),
        ("gy", ctypes.c_int),
        ("gz", ctypes.c_int),
        ("dx", ctypes.c_float),
        ("rdx", ctypes.c_float),
        ("velocity_emission_rate", ctypes.c_float),
        ("density_emission_rate", ctypes.c_float),
        ("emission_radius", ctypes.c_float),
        ("enable_additional_solids", ctypes.c_bool),
        ("render_solids", ctypes.c_bool),
        ("render_shadows", ctypes.c_bool),
        ("render_precise", ctypes.c_bool),
        ("render_channel", ctypes.c_int32),
        ("aa_samples", ctypes.c_float),
        ("ambient", ctypes.c_float),
        ("background_color", Float3),
        ("floor_color", Float3),
        ("smoke_color", Float3),
        ("scene_color", Float3),
    ]


assert ctypes.sizeof(SimParams) == 100


def default_params() -> SimParams:
    """SimParams::defaultParams() + rdx (src/data/sim_params.cpp:4-34, src/main.cpp:17)."""
    p = SimParams()
    p.gx = p.gy = p.gz = 128
    p.dx = np.float32(10000.0) / np.float32(128)
    p.dt = 3.0
    p.velocity_emission_rate = 150.0
    p.density_emission_rate = 0.002
    p.emission_radius = 750.0
    p.enable_additional_solids = False
    p.render_solids = True
    p.render_shadows = True
    p.render_precise = True
    p.render_channel = 4  # RenderChannel::Resolution
    p.aa_samples = 1.0
    p.ambient = 0.3
    p.background_color = Float3(0.0, 0.0, 0.0)
    p.floor_color = Float3(*(np.float32([178.0, 158.0, 135.0]) / np.float32(255.0)))
    p.smoke_color = Float3(0.9, 0.9, 0.9)
    p.scene_color = Float3(*(np.float32([107.0, 163.0, 204.0]) / np.float32(255.0)))
    p.rdx = np.float32(1.0) / np.float32(p.dx)
    return p


def scene_params(gx: int, gy: int = None, gz: int = None, solids: bool = False) -> SimParams:
    """Scene preset in the style of src/data/scenes.h: dx = 10000/gx, rdx = 1/dx."""
    p = default_params()
    p.gx = gx
    p.gy = gx if gy is None else gy
    p.gz = gx if gz is None else gz
    p.dx = np.float32(10000.0) / np.float32(gx)
    p.rdx = np.float32(1.0) / np.float32(p.dx)
    p.enable_additional_solids = solids
    return p


class Options(ctypes.Structure):
    """``struct dcg_options`` (include/dcgrid_b200.h): creation-time options, every field 0 = default."""
    _fields_ = [("struct_size", ctypes.c_uint32)] + [(n, ctypes.c_int32) for n in (
        "jacobi", "jacobi_ctas_per_sm", "no_snake", "advect", "advect_slot_order", "advect_no_fuse", "advect_min_blocks",
        "advect_ctas_per_sm", "stencil", "stencil_ctas_per_sm", "apply_min_blocks", "coarse_in_gmem", "zero_all", "no_resort",
        "resort_every")] + [("shard_unit", ctypes.c_uint32), ("no_pdl", ctypes.c_int32), ("host_selection", ctypes.c_int32),
                            ("jacobi_max_ctas", ctypes.c_int32), ("experiment", ctypes.c_int32), ("reserved", ctypes.c_int32 * 12)]


def make_options(options=None) -> Options:
    """dict (or Options, or None) -> Options with struct_size filled in."""
    if isinstance(options, Options):
        o = Options.from_buffer_copy(options)
    else:
        o = Options()
        for k, v in (options or {}).items():
            if not hasattr(o, k):
                raise KeyError(f"unknown dcg_options field {k!r}")
            setattr(o, k, int(v))
    o.struct_size = ctypes.sizeof(Options)
    return o


class ExtParams(ctypes.Structure):
    """``struct dcg_ext_params`` (include/dcgrid_b200.h): extensions beyond the reference snapshot; all zero = off."""
    _fields_ = [("struct_size", ctypes.c_uint32)] + [(n, ctypes.c_int32) for n in ("score_mode", "advection", "sources", "terrain", "selection")] + [
        (n, ctypes.c_float) for n in (
            "buoyancy", "vapor_buoyancy", "smoke_weight", "ambient_temperature", "ambient_lapse", "adiabatic_lapse", "vorticity_confinement",
            "saturation_base", "saturation_slope", "condensation_rate", "latent_heat", "temperature_emission", "vapor_emission", "ambient_vapor",
            "terrain_height", "terrain_wavelength")] + [("reserved", ctypes.c_int32 * 10)]


def make_ext(**kw) -> ExtParams:
    """dcg_default_ext_params() (plausible coefficients, every switch off) with the given fields overridden."""
    e = ExtParams()
    from . import _lib

    _lib.load().dcg_default_ext_params(ctypes.byref(e))
    for k, v in kw.items():
        if not hasattr(e, k):
            raise KeyError(f"unknown dcg_ext_params field {k!r}")
        setattr(e, k, v)
    e.struct_size = ctypes.sizeof(ExtParams)
    return e
