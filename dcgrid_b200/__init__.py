"""dcgrid_b200 — B200-native drop-in for the per-timestep fluid solve of wouterraateland/dcgrid.

Only what the hot path needs: ``csrc/`` (hand-written sm_100a CUDA + the C ABI of
include/dcgrid_b200.h) and this thin host-side mirror of the reference's FluidSimulation
interface.  Importing the package does not load CUDA; constructing a simulation does and
fails loudly when the extension or a GPU is missing.
"""
from .params import ExtParams, SimParams, default_params, make_ext, scene_params  # noqa: F401
from .simulation import (  # noqa: F401
    DcgError,
    FluidSimulation,
    FluidSimulationDCGrid,
    FluidSimulationUniform,
    fnv1a64,
)
from .sharding import FluidSimulationDCGridSharded, FluidSimulationUniformSharded  # noqa: F401,E402
