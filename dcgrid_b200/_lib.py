"""ctypes loader for libdcgrid_b200.so (the C ABI declared in include/dcgrid_b200.h).

There is no CPU fallback: if the CUDA library is missing it is built (nvcc), and if that
fails the import raises.
"""
import ctypes
import os

from .params import ExtParams, Options, SimParams

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libdcgrid_b200.so")

# every symbol include/dcgrid_b200.h declares: name -> (restype, argtypes)
_vp = ctypes.c_void_p
_P = ctypes.POINTER(SimParams)
_O = ctypes.POINTER(Options)
_E = ctypes.POINTER(ExtParams)
_u64 = ctypes.c_uint64
_int = ctypes.c_int
SYMBOLS = {
    "dcg_default_params": (_int, [_P]),
    "dcg_create_uniform": (_int, [_P, _int, ctypes.POINTER(_vp)]),
    "dcg_create_dcgrid": (_int, [_P, _u64, _int, ctypes.POINTER(_vp)]),
    "dcg_default_options": (_int, [_O]),
    "dcg_create_uniform_opt": (_int, [_P, _int, _O, ctypes.POINTER(_vp)]),
    "dcg_create_dcgrid_opt": (_int, [_P, _u64, _int, _O, ctypes.POINTER(_vp)]),
    "dcg_create_dcgrid_sharded_opt": (_int, [_P, ctypes.c_uint64, _int, _int, _int, _int, _O, ctypes.POINTER(_vp)]),
    "dcg_destroy": (_int, [_vp]),
    "dcg_set_params": (_int, [_vp, _P]),
    "dcg_get_params": (_int, [_vp, _P]),
    "dcg_init": (_int, [_vp]),
    "dcg_reset": (_int, [_vp]),
    "dcg_adapt_topology": (_int, [_vp]),
    "dcg_advect_velocity": (_int, [_vp]),
    "dcg_project": (_int, [_vp]),
    "dcg_project_local": (_int, [_vp]),
    "dcg_advect_density": (_int, [_vp]),
    "dcg_render": (_int, [_vp]),
    "dcg_debug_stats": (_int, [_vp, ctypes.POINTER(ctypes.c_float)]),
    "dcg_step": (_int, [_vp, _int]),
    "dcg_synchronize": (_int, [_vp]),
    "dcg_set_jacobi_schedule": (_int, [_vp, _int, _int, _int]),
    "dcg_total_density": (_int, [_vp, ctypes.POINTER(ctypes.c_double)]),
    "dcg_is_dcgrid": (_int, [_vp]),
    "dcg_num_cells": (_u64, [_vp]),
    "dcg_max_num_blocks": (_u64, [_vp]),
    "dcg_num_levels": (_int, [_vp]),
    "dcg_sparse_levels": (_int, [_vp]),
    "dcg_get_field": (_int, [_vp, _int, _int, _vp, _u64]),
    "dcg_get_level_table": (_int, [_vp, _vp, _vp, _vp, _vp]),
    "dcg_get_topology": (_int, [_vp, _vp, _vp, _vp, _vp, _vp]),
    "dcg_lookup_blocks": (_int, [_vp, _vp, _u64, _vp, _vp]),
    "dcg_get_counters": (_int, [_vp, _vp]),
    "dcg_get_info": (_int, [_vp, ctypes.c_char_p, ctypes.POINTER(ctypes.c_double)]),
    "dcg_last_step_ms": (_int, [_vp, ctypes.POINTER(ctypes.c_float)]),
    "dcg_algorithmic_bytes": (_int, [_vp, ctypes.POINTER(ctypes.c_double), ctypes.POINTER(_u64)]),
    "dcg_bench_stage": (_int, [_vp, ctypes.c_char_p, _int, _int, ctypes.POINTER(ctypes.c_float), ctypes.POINTER(ctypes.c_double)]),
    "dcg_create_uniform_sharded": (_int, [_P, _int, _int, _int, _int, ctypes.POINTER(_vp)]),
    "dcg_create_dcgrid_sharded": (_int, [_P, ctypes.c_uint64, _int, _int, _int, _int, ctypes.POINTER(_vp)]),
    "dcg_shard_handle_bytes": (_u64, []),
    "dcg_shard_export_handle": (_int, [_vp, _vp, _u64]),
    "dcg_shard_import_handles": (_int, [_vp, _vp, _int]),
    "dcg_fnv1a64": (_u64, [_vp, _u64, _u64]),
    "dcg_default_ext_params": (_int, [_E]),
    "dcg_set_ext_params": (_int, [_vp, _E]),
    "dcg_get_ext_params": (_int, [_vp, _E]),
    "dcg_apply_sources": (_int, [_vp]),
    "dcg_sample_field": (_int, [_vp, _int, _int, _vp, _u64, _vp]),
    "dcg_save_state": (_int, [_vp, ctypes.c_char_p]),
    "dcg_load_state": (_int, [_vp, ctypes.c_char_p]),
    "dcg_last_error": (ctypes.c_char_p, [_vp]),
    "dcg_version": (ctypes.c_char_p, []),
}

_lib = None


def load(build_if_missing=True):
    """Loads the shared library and types every exported symbol.  Raises if it is absent."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        if not build_if_missing:
            raise ImportError(f"{LIB_PATH} is missing; run `python -m dcgrid_b200.build` (no CPU fallback exists)")
        from . import build as _build

        _build.build()
    L = ctypes.CDLL(LIB_PATH)
    for name, (res, args) in SYMBOLS.items():
        fn = getattr(L, name)  # AttributeError if the .so does not export a declared symbol
        fn.restype = res
        fn.argtypes = args
    _lib = L
    return L
