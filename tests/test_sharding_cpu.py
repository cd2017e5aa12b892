"""CPU tests of the multi-GPU host logic: slab / mip-plane ownership arithmetic and the bootstrap
collectives (handle all-gather, partial-sum all-reduce) over gloo with world_size 2."""
import os
import socket

import numpy as np
import pytest
import torch.distributed as dist
import torch.multiprocessing as mp

from dcgrid_b200 import sharding


def test_slabs_partition_every_level():
    for gz, world in ((64, 2), (64, 8), (48, 3), (1024, 8), (32, 8)):
        levels = 1
        while gz % (1 << levels) == 0 and (1 << levels) * 4 <= gz:
            levels += 1
        assert sharding.slab_range(gz, world, 0)[0] == 0 and sharding.slab_range(gz, world, world - 1)[1] == gz
        for l in range(levels):
            planes = gz >> l
            owned = np.full(planes, -1)
            for r in range(world):
                z0, z1 = sharding.level_planes(gz, world, r, l)
                assert (owned[z0:z1] == -1).all()
                owned[z0:z1] = r
                for z in range(z0, z1):
                    assert sharding.plane_owner(gz, world, l, z) == r
            assert (owned >= 0).all(), (gz, world, l)   # every plane has exactly one owner
            # a coarse plane lives where its first fine plane lives
            if l > 0:
                for z in range(planes):
                    assert owned[z] == sharding.plane_owner(gz, world, l - 1, 2 * z) or (gz // world) % (1 << l) != 0


def test_slab_range_rejects_ragged():
    with pytest.raises(ValueError):
        sharding.slab_range(30, 4, 0)


def test_halo_traffic_is_two_planes_for_interior_ranks():
    assert sharding.halo_bytes_per_sweep(1024, 1024, 1024, 8, 3) == 2 * 1024 * 1024 * 4
    assert sharding.halo_bytes_per_sweep(1024, 1024, 1024, 8, 0) == 1024 * 1024 * 4
    assert sharding.halo_bytes_per_sweep(64, 64, 64, 1, 0) == 0


def test_uniform_pieces_follow_cell_ownership():
    """Placement of the sharded dense solver's fields (uniform_piece_runs mirrors UniformSim::build_piece_runs): a rank's
    level-0 slab lies in its own physical pieces, and so do its planes of every pyramid level big enough for a granule —
    with equal byte shares 14 % of a rank's level-0 pressure cells were remote (2 GPUs ran at 1.36x one)."""
    gran = 2 << 20
    for (gx, gy, gz), world in (((1024, 1024, 1024), 2), ((1024, 1024, 1024), 8), ((512, 512, 512), 4), ((256, 128, 512), 4)):
        slab = gz // world
        for field, item in (("vw", 16), ("q", 4)):
            runs = sharding.uniform_piece_runs(gx, gy, gz, world, gran, field)
            assert runs[0][0] == 0 and all(a[1] == b[0] for a, b in zip(runs, runs[1:]))     # contiguous cover
            assert [r[2] for r in runs] == sorted(r[2] for r in runs)                          # one run per rank, in order
            slab_bytes = gx * gy * slab * item
            if slab_bytes % gran == 0:
                assert [(r[0], r[1]) for r in runs] == [(k * slab_bytes, (k + 1) * slab_bytes) for k in range(world)]
        runs = sharding.uniform_piece_runs(gx, gy, gz, world, gran, "pyramid")
        assert runs[0][0] == 0 and all(a[1] == b[0] for a, b in zip(runs, runs[1:]))
        # level 0 of the pyramid: every rank's slab in its own pieces (slab bytes are granule multiples at these sizes)
        n0 = gx * gy * gz
        for k in range(world):
            lo, hi = k * gx * gy * slab * 4, (k + 1) * gx * gy * slab * 4
            if (gx * gy * slab * 4) % gran:
                continue
            covered = sum(min(hi, r[1]) - max(lo, r[0]) for r in runs if r[2] == k and r[1] > lo and r[0] < hi)
            assert covered == hi - lo, ((gx, gy, gz), world, k)
        # level 1 starts a new round of owners right after level 0
        after_l0 = [r for r in runs if r[0] >= n0 * 4]
        if after_l0 and (gx * gy * slab * 4 // 8) % gran == 0:
            assert after_l0[0][2] == 0
    assert sharding.uniform_piece_runs(64, 64, 64, 1, gran, "pyramid") == [(0, gran * ((sum((64 >> l) ** 3 for l in range(7)) * 4 + gran - 1) // gran), 0)]


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        blob = bytes([rank + 1]) * 64                      # stands in for a cudaIpcMemHandle_t
        got = sharding.all_gather_bytes(blob, dist)
        total = sharding.all_reduce_sum(1.5 * (rank + 1), dist)
        q.put((rank, [g[0] for g in got], [len(g) for g in got], total, sharding.slab_range(64, world, rank)))
    finally:
        dist.destroy_process_group()


def test_bootstrap_collectives_gloo_world2():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, firsts, lens, total, zr in res:
        assert firsts == [1, 2] and lens == [64, 64]       # rank order, fixed size
        assert total == 4.5
        assert zr == (32 * rank, 32 * rank + 32)


def test_dcgrid_unit_ownership_is_contiguous_and_balanced_per_level():
    """Ownership table of the slab-decomposed adaptive solver (dcgrid_unit_owner mirrors
    DCGridSim::setup_sharding): inside a level the owner never decreases with the slot, and every rank gets
    its share of a level that spans enough units."""
    from dcgrid_b200.sharding import dcgrid_unit_owner

    # C3's level table: 512^3, M = 524,288
    max_blocks = [243420, 243419, 32768, 4096, 512, 64, 8, 1]
    offsets = np.concatenate([[0], np.cumsum(max_blocks)[:-1]]).tolist()
    M = 524288
    for world, unit in ((2, 8192), (4, 8192), (8, 8192), (3, 16), (8, 16)):
        own = dcgrid_unit_owner(M, offsets, max_blocks, world, unit)
        assert own.shape[0] == (M + unit - 1) // unit and own.max() <= world - 1
        mids = np.minimum(np.arange(own.shape[0]) * unit + unit // 2, M - 1)
        for off, mx in zip(offsets, max_blocks):
            sel = (mids >= off) & (mids < off + mx)
            o = own[sel]
            if o.size == 0:
                continue
            assert np.all(np.diff(o.astype(int)) >= 0), "owners are contiguous runs inside a level"
            if o.size >= 4 * world:
                counts = np.bincount(o, minlength=world)
                assert counts.min() >= o.size // world - 1 and counts.max() <= o.size // world + 2
    assert np.all(dcgrid_unit_owner(M, offsets, max_blocks, 1, 8192) == 0)


def test_descriptor_exchange_between_two_processes(tmp_path):
    """The slab-decomposed DCGrid solver passes its GPU allocations between ranks as POSIX file descriptors over
    abstract AF_UNIX sockets (dcgrid_b200/csrc/shard_vmm.h: FdServer / fetch_fds).  That plumbing is GPU-free:
    two processes exchange three descriptors each (pipes carrying a rank-specific message) and check them."""
    import os
    import shutil
    import subprocess

    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    cuda_inc = os.environ.get("CUDA_HOME", "/usr/local/cuda") + "/include"
    if shutil.which("g++") is None or not os.path.exists(os.path.join(cuda_inc, "cuda.h")):
        import pytest

        pytest.skip("needs g++ and the CUDA headers (types only; nothing is linked)")
    exe = str(tmp_path / "fd_exchange_test")
    subprocess.check_call(["g++", "-O1", "-std=c++17", "-I" + cuda_inc, os.path.join(root, "tests", "native", "fd_exchange_test.cpp"), "-o", exe,
                           "-pthread"])
    r = subprocess.run([exe], capture_output=True, text=True, timeout=120)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "rank0 rc=0 rank1 rc=0" in r.stdout
