"""ctypes binding of oracle/liboracle.so — the CPU restatement used as the checker.

Test infrastructure only: nothing under dcgrid_b200/ imports this.
"""
import ctypes
import os
import subprocess

import numpy as np

from dcgrid_b200.params import ExtParams, SimParams

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "oracle", "liboracle.so")

FIELD = {"density": 0, "velocity": 1, "fluidity": 2, "pressure": 3, "divergence": 4, "t_pressure": 5, "temperature": 6, "vapor": 7, "vorticity": 8}


def build():
    src = os.path.join(ROOT, "oracle", "dcgrid_oracle.cpp")
    if not os.path.exists(LIB) or os.path.getmtime(LIB) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", os.path.join(ROOT, "oracle"), "liboracle"], stdout=subprocess.DEVNULL)
    return LIB


_lib = None


def lib():
    global _lib
    if _lib is None:
        L = ctypes.CDLL(build())
        P = ctypes.POINTER(SimParams)
        vp = ctypes.c_void_p
        L.orc_create_uniform.restype = vp
        L.orc_create_uniform.argtypes = [P]
        L.orc_create_dcgrid.restype = vp
        L.orc_create_dcgrid.argtypes = [P, ctypes.c_uint64]
        L.orc_destroy.argtypes = [vp]
        L.orc_set_params.argtypes = [vp, P]
        L.orc_set_jacobi_schedule.argtypes = [vp, ctypes.c_int, ctypes.c_int, ctypes.c_int]
        for n in ("init", "reset", "adapt_topology", "advect_velocity", "project", "project_local", "advect_density"):
            getattr(L, "orc_" + n).argtypes = [vp]
        L.orc_step.argtypes = [vp, ctypes.c_int]
        L.orc_debug_stats.restype = ctypes.c_float
        L.orc_debug_stats.argtypes = [vp]
        L.orc_num_cells.restype = ctypes.c_uint64
        L.orc_num_cells.argtypes = [vp]
        L.orc_get_field.argtypes = [vp, ctypes.c_int, vp]
        L.orc_num_levels.argtypes = [vp]
        L.orc_sparse_levels.argtypes = [vp]
        L.orc_get_level_table.argtypes = [vp, vp, vp, vp, vp]
        L.orc_get_topology.argtypes = [vp, vp, vp, vp, vp, vp]
        L.orc_lookup_blocks.argtypes = [vp, vp, ctypes.c_uint64, vp, vp]
        L.orc_get_counters.argtypes = [vp, vp]
        L.orc_get_move_limits.argtypes = [vp, vp]
        L.orc_set_ext_params.argtypes = [vp, ctypes.POINTER(ExtParams)]
        L.orc_apply_sources.argtypes = [vp]
        L.orc_sample_field.argtypes = [vp, ctypes.c_int, ctypes.c_int, vp, ctypes.c_uint64, vp]
        _lib = L
    return _lib


def _ptr(a):
    return a.ctypes.data_as(ctypes.c_void_p)


class Oracle:
    """Mirror of the FluidSimulation interface (src/fluid_simulation.h:4-27) on the CPU oracle."""

    def __init__(self, params: SimParams, max_num_blocks: int = 0):
        self.L = lib()
        self.params = params
        self.is_dcgrid = max_num_blocks > 0
        if self.is_dcgrid:
            self.h = self.L.orc_create_dcgrid(ctypes.byref(params), max_num_blocks)
        else:
            self.h = self.L.orc_create_uniform(ctypes.byref(params))
        if not self.h:
            raise RuntimeError("oracle: pool too small")
        self.max_num_blocks = max_num_blocks

    def close(self):
        if self.h:
            self.L.orc_destroy(self.h)
            self.h = None

    def __del__(self):
        self.close()

    def set_jacobi_schedule(self, coarse, level, local):
        self.L.orc_set_jacobi_schedule(self.h, coarse, level, local)

    def set_params(self, params: SimParams):
        self.params = params
        self.L.orc_set_params(self.h, ctypes.byref(params))

    def set_ext(self, ext: ExtParams):
        """extensions: this oracle is their specification"""
        self.ext = ext
        self.L.orc_set_ext_params(self.h, ctypes.byref(ext))

    def apply_sources(self):
        self.L.orc_apply_sources(self.h)

    def sample_field(self, name, positions, precise=False):
        positions = np.ascontiguousarray(positions, dtype=np.float32).reshape(-1, 3)
        comps = 3 if name in ("velocity", "vorticity") else 1
        out = np.empty(positions.shape[0] * comps, dtype=np.float32)
        rc = self.L.orc_sample_field(self.h, FIELD[name], 1 if precise else 0, _ptr(positions), positions.shape[0], _ptr(out))
        assert rc == 0
        return out.reshape(-1, 3) if comps == 3 else out

    def reset(self):
        self.L.orc_reset(self.h)

    def adapt_topology(self):
        self.L.orc_adapt_topology(self.h)

    def advect_velocity(self):
        self.L.orc_advect_velocity(self.h)

    def project(self):
        self.L.orc_project(self.h)

    def project_local(self):
        self.L.orc_project_local(self.h)

    def advect_density(self):
        self.L.orc_advect_density(self.h)

    def step(self, n=1):
        self.L.orc_step(self.h, n)

    def debug_stats(self):
        return float(self.L.orc_debug_stats(self.h))

    @property
    def num_cells(self):
        return int(self.L.orc_num_cells(self.h))

    @property
    def levels(self):
        return int(self.L.orc_num_levels(self.h))

    @property
    def sparse_levels(self):
        return int(self.L.orc_sparse_levels(self.h))

    def field(self, name):
        vec = name in ("velocity", "vorticity")
        n = self.num_cells * (3 if vec else 1)
        out = np.empty(n, dtype=np.float32)
        rc = self.L.orc_get_field(self.h, FIELD[name], _ptr(out))
        assert rc == 0
        return out.reshape(-1, 3) if vec else out

    def level_table(self):
        L = self.levels
        arrs = [np.zeros(L, dtype=np.uint64) for _ in range(4)]
        self.L.orc_get_level_table(self.h, *[_ptr(a) for a in arrs])
        return dict(zip(("max_blocks", "full_blocks", "loads", "offsets"), arrs))

    def topology(self, with_apron=True):
        M = self.max_num_blocks
        pos = np.zeros((M, 3), dtype=np.int32)
        lvl = np.zeros(M, dtype=np.uint8)
        parent = np.zeros(M, dtype=np.uint64)
        child = np.zeros((M, 8), dtype=np.uint64)
        apron = np.zeros((M, 216), dtype=np.uint64) if with_apron else None
        self.L.orc_get_topology(self.h, _ptr(pos), _ptr(lvl), _ptr(parent), _ptr(child),
                                _ptr(apron) if with_apron else None)
        return dict(pos=pos, level=lvl, parent=parent, child=child, apron=apron)

    def lookup_blocks(self, positions):
        positions = np.ascontiguousarray(positions, dtype=np.int32)
        n = positions.shape[0]
        slot = np.zeros(n, dtype=np.uint64)
        lvl = np.zeros(n, dtype=np.uint8)
        self.L.orc_lookup_blocks(self.h, _ptr(positions), n, _ptr(slot), _ptr(lvl))
        return slot, lvl

    def counters(self):
        out = np.zeros(8, dtype=np.uint64)
        self.L.orc_get_counters(self.h, _ptr(out))
        return out

    def move_limits(self):
        out = np.zeros(self.levels, dtype=np.uint64)
        self.L.orc_get_move_limits(self.h, _ptr(out))
        return out
