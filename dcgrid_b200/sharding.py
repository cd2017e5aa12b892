"""Host-side logic of the z-slab decomposition (multi-GPU uniform grid).

The reference is single-GPU (src/main.cpp:46-48); this is new design, see DESIGN.md §6.  One process per
GPU (torchrun); ``torch.distributed`` is only the bootstrap channel (it carries 64-byte handles once, and
scalar reductions of residual / smoke totals).  The data path never goes through it: every field is one virtual
address range stitched from the ranks' physical pieces, the single-GPU kernels run on each rank's planes and read
neighbour slabs in place over NVLink (dcgrid_b200/csrc/uniform.cu, "slab decomposition").

Everything here is pure host arithmetic + collectives, so it runs (and is tested) on CPU with gloo.
"""
import ctypes

import numpy as np

from . import _lib
from .params import SimParams, make_options
from .simulation import FluidSimulation, FluidSimulationDCGrid


def slab_range(gz, world, rank):
    """Level-0 z-planes [z0, z1) owned by ``rank``: equal contiguous slabs (gz % world == 0)."""
    if gz % world:
        raise ValueError(f"gz={gz} is not a multiple of world={world}")
    s = gz // world
    return rank * s, (rank + 1) * s


def first_plane(slab, level, rank):
    """First plane of mip level ``level`` owned by ``rank``: a coarse plane belongs to the owner of its
    first fine plane (same rule as UniformSim::zfirst() in uniform.cu)."""
    return (rank * slab + (1 << level) - 1) >> level


def level_planes(gz, world, rank, level):
    s = gz // world
    return first_plane(s, level, rank), first_plane(s, level, rank + 1)


def plane_owner(gz, world, level, z):
    return min((z << level) // (gz // world), world - 1)


def halo_bytes_per_sweep(gx, gy, gz, world, rank, level=0):
    """Bytes a rank reads from its peers in one Jacobi sweep of ``level`` (its z-1 and z+1 face planes)."""
    z0, z1 = level_planes(gz, world, rank, level)
    if z1 <= z0:
        return 0
    faces = (1 if z0 > 0 else 0) + (1 if z1 < (gz >> level) else 0)
    return faces * (gx >> level) * (gy >> level) * 4


def uniform_piece_runs(gx, gy, gz, world, gran, field):
    """Physical placement of one field of the sharded dense solver (same arithmetic as UniformSim::build_piece_runs in
    csrc/uniform.cu): the field's byte range in allocation granules of ``gran`` bytes, each granule on the rank that owns
    the cell in its middle; returns the maximal runs [(byte0, byte1, owner), ...].  ``field``: "vw" (float4 per level-0
    cell), "q" (float per level-0 cell) or "pyramid" (float per cell of every mip level, levels concatenated)."""
    slab = gz // world
    n0 = gx * gy * gz
    if field == "pyramid":
        offs, s, cells = [], 1, 0
        while gx % s == 0 and gy % s == 0 and gz % s == 0:   # grid_math.cuh:24-30 (mipmapCells)
            offs.append(cells)
            cells += n0 // (s * s * s)
            s *= 2
        min_dim, levels, cell = min(gx, gy, gz), 1, 2
        while gx % cell == 0 and gy % cell == 0 and gz % cell == 0 and cell * 4 <= min_dim:   # fluid_simulation_uniform.cu:8-17
            levels += 1
            cell *= 2
        item, total_cells = 4, cells
    else:
        item, total_cells = (16 if field == "vw" else 4), n0

    def owner(byte):
        c = min(total_cells - 1, byte // item)
        if field != "pyramid":
            return min(world - 1, c // (gx * gy) // slab)
        l = 0
        while l + 1 < levels and c >= offs[l + 1]:
            l += 1
        w, h, d = gx >> l, gy >> l, gz >> l
        if c >= offs[l] + w * h * d:
            return world - 1
        z = (c - offs[l]) // (w * h)
        return min(world - 1, (z << l) // slab)

    runs = []
    ngran = (total_cells * item + gran - 1) // gran
    for g in range(ngran):
        o = owner(g * gran + gran // 2)
        if runs and runs[-1][2] == o:
            runs[-1] = (runs[-1][0], (g + 1) * gran, o)
        else:
            runs.append((g * gran, (g + 1) * gran, o))
    return runs


def all_gather_bytes(blob: bytes, dist=None):
    """All-gathers one fixed-size byte blob per rank, in rank order.  ``dist`` = torch.distributed (any
    backend; gloo on CPU in the tests, nccl on the GPU box) or None for a single process."""
    if dist is None or not dist.is_initialized() or dist.get_world_size() == 1:
        return [bytes(blob)]
    import torch

    dev = "cuda" if dist.get_backend() == "nccl" else "cpu"
    mine = torch.frombuffer(bytearray(blob), dtype=torch.uint8).to(dev)
    out = [torch.empty_like(mine) for _ in range(dist.get_world_size())]
    dist.all_gather(out, mine)
    return [bytes(t.cpu().numpy().tobytes()) for t in out]


def all_reduce_sum(value: float, dist=None):
    """Sum of a per-rank partial (smoke total, residual) over all ranks, in float64."""
    if dist is None or not dist.is_initialized() or dist.get_world_size() == 1:
        return float(value)
    import torch

    dev = "cuda" if dist.get_backend() == "nccl" else "cpu"
    t = torch.tensor([value], dtype=torch.float64, device=dev)
    dist.all_reduce(t)
    return float(t.item())


class FluidSimulationUniformSharded(FluidSimulation):
    """Ranks [rank, rank+nlocal) of a ``world``-way z-slab decomposition of FluidSimulationUniform(size).

    nlocal == world: every rank lives in this object on one device (tests the decomposition on one GPU).
    nlocal == 1: one rank per process; pass ``dist`` (an initialised torch.distributed) to exchange the IPC
    handles.  After construction all ranks must issue the same sequence of solver calls."""

    def __init__(self, size, params: SimParams, world, rank=0, nlocal=None, device=0, dist=None):
        super().__init__()
        p = SimParams.from_buffer_copy(params)
        p.gx, p.gy, p.gz = size
        self.params, self.world, self.rank = p, int(world), int(rank)
        self.nlocal = self.world if nlocal is None else int(nlocal)
        self.dist = dist
        self._check(self._L.dcg_create_uniform_sharded(ctypes.byref(p), device, self.rank, self.world, self.nlocal, ctypes.byref(self._h)))
        if self.nlocal != self.world:
            n = int(self._L.dcg_shard_handle_bytes())
            buf = ctypes.create_string_buffer(n)
            self._check(self._L.dcg_shard_export_handle(self._h, buf, n))
            blobs = all_gather_bytes(buf.raw, dist)
            if len(blobs) != self.world:
                raise ValueError(f"expected {self.world} IPC handles, got {len(blobs)}")
            allb = b"".join(blobs)
            self._check(self._L.dcg_shard_import_handles(self._h, allb, self.world))

    @property
    def z_range(self):
        s = self.params.gz // self.world
        return self.rank * s, (self.rank + self.nlocal) * s

    def totalDensity(self):
        """Global smoke total: local partial summed over ranks."""
        return all_reduce_sum(FluidSimulation.totalDensity(self), self.dist if self.nlocal != self.world else None)


def dcgrid_unit_owner(max_num_blocks, offsets, max_blocks, world, unit):
    """Owner rank of every ``unit``-slot piece of the block pool: each level's slot range is shared equally and
    contiguously (same arithmetic as DCGridSim::setup_sharding in csrc/dcgrid.cu)."""
    M = int(max_num_blocks)
    n = (M + unit - 1) // unit
    owner = np.zeros(n, dtype=np.uint8)
    for u in range(n):
        mid = min(u * unit + unit // 2, M - 1)
        for off, mx in zip(offsets, max_blocks):
            if off <= mid < off + mx:
                owner[u] = min(world - 1, (mid - off) * world // mx) if world > 1 else 0
                break
    return owner


class FluidSimulationDCGridSharded(FluidSimulationDCGrid):
    """Ranks [rank, rank+nlocal) of a ``world``-way slab decomposition of FluidSimulationDCGrid(size, maxNumBlocks).

    nlocal == world: every rank lives in this object on one device (tests the decomposition on one GPU).
    nlocal == 1: one rank per process; pass ``dist`` (an initialised torch.distributed) to exchange the arena
    handles.  After construction all ranks must issue the same sequence of solver calls."""

    def __init__(self, size, maxNumBlocks, params: SimParams, world, rank=0, nlocal=None, device=0, dist=None, options=None):
        FluidSimulation.__init__(self)
        p = SimParams.from_buffer_copy(params)
        p.gx, p.gy, p.gz = size
        self.params, self.world, self.rank = p, int(world), int(rank)
        self.nlocal = self.world if nlocal is None else int(nlocal)
        self.dist = dist
        self.maxNumBlocks = int(maxNumBlocks)
        o = make_options(options)
        self._check(self._L.dcg_create_dcgrid_sharded_opt(ctypes.byref(p), self.maxNumBlocks, device, self.rank, self.world, self.nlocal,
                                                          ctypes.byref(o), ctypes.byref(self._h)))
        if self.nlocal != self.world:
            n = int(self._L.dcg_shard_handle_bytes())
            buf = ctypes.create_string_buffer(n)
            self._check(self._L.dcg_shard_export_handle(self._h, buf, n))
            blobs = all_gather_bytes(buf.raw, dist)
            if len(blobs) != self.world:
                raise ValueError(f"expected {self.world} arena handles, got {len(blobs)}")
            self._check(self._L.dcg_shard_import_handles(self._h, b"".join(blobs), self.world))

    def totalDensity(self):
        """Global smoke total: the partial over the cells this instance owns, summed over ranks."""
        return all_reduce_sum(FluidSimulation.totalDensity(self), self.dist if self.nlocal != self.world else None)
