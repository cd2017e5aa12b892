"""Multi-process check of the slab-decomposed adaptive solver (one process per GPU, fields in one virtual range
stitched from every GPU's arena, flag barriers over NVLink).

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
        tests/mgpu_dcgrid_check.py [--size 128] [--blocks 65536] [--steps 8] [--bench-size 512] [--bench-blocks 524288]

Every rank runs (a) its share of the N-rank decomposition and (b) the plain single-GPU solver on the whole pool,
and compares the whole pool bit for bit (any rank can read every cell).  Then the sharded solver is timed on a
larger scene (device time, max over ranks).  Not collected by pytest: needs >= 2 GPUs (gpurun --gpus 2)."""
import argparse
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--size", type=int, default=128)
    ap.add_argument("--blocks", type=int, default=65536)
    ap.add_argument("--steps", type=int, default=8)
    ap.add_argument("--bench-size", type=int, default=512)
    ap.add_argument("--bench-blocks", type=int, default=524288)
    ap.add_argument("--bench-steps", type=int, default=20)
    ap.add_argument("--bench-warm", type=int, default=150)
    args = ap.parse_args()
    import torch
    import torch.distributed as dist

    from dcgrid_b200 import FluidSimulationDCGrid, FluidSimulationDCGridSharded, scene_params

    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    ok = True
    d, M = args.size, args.blocks
    p = scene_params(d, solids=True)
    sh = FluidSimulationDCGridSharded((d, d, d), M, p, world, rank=rank, nlocal=1, device=local, dist=dist)
    one = FluidSimulationDCGrid((d, d, d), M, p, device=local)

    def compare(fields, where):
        nonlocal ok
        ta, tb = sh.topology(), one.topology()
        for k in ("level", "pos", "parent", "child", "apron"):
            if not np.array_equal(ta[k], tb[k]):
                ok = False
                print(f"[rank {rank}] MISMATCH topology {k} {where}", flush=True)
        for f in fields:
            a, b = sh.field(f), one.field(f)
            if not np.array_equal(np.ascontiguousarray(a).view(np.uint32), np.ascontiguousarray(b).view(np.uint32)):
                ok = False
                print(f"[rank {rank}] MISMATCH {f} {where}: max abs {np.abs(a - b).max()}", flush=True)

    compare(("density", "velocity", "fluidity"), "after reset")
    for s in range(args.steps):
        for sim in (sh, one):
            sim.advectVelocity(); sim.adaptTopology(); sim.project()
        sh.synchronize(); dist.barrier()
        if s == args.steps - 1:
            compare(("pressure", "t_pressure", "divergence", "velocity"), f"after project, step {s}")
        dist.barrier()
        for sim in (sh, one):
            sim.advectDensity()
        sh.synchronize(); dist.barrier()
    compare(("density", "velocity"), "at the end")
    dist.barrier()
    sh.step(6); one.step(6)  # dcg_step path (graphs once the topology is at its fixed point)
    sh.synchronize(); dist.barrier()
    compare(("density", "velocity"), "after step(6)")
    tot_s, tot_1 = sh.totalDensity(), one.totalDensity()
    ok = ok and abs(tot_s - tot_1) <= 1e-9 * max(1.0, abs(tot_1))
    barriers = int(sh.counters()[7]) if False else None
    dist.barrier()
    sh.close(); one.close()
    flag = torch.tensor([1 if ok else 0], device="cuda")
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    out = {"check": "mgpu_dcgrid", "n_gpus": world, "bit_exact_vs_single_gpu": bool(flag.item()), "parity_size": d, "parity_blocks": M}

    if args.bench_size > 0:
        d, M = args.bench_size, args.bench_blocks
        p = scene_params(d, solids=True)
        sh = FluidSimulationDCGridSharded((d, d, d), M, p, world, rank=rank, nlocal=1, device=local, dist=dist)
        sh.step(args.bench_warm)
        dist.barrier(); torch.cuda.synchronize()
        sh.step(args.bench_steps, sync=True)
        t = torch.tensor([sh.lastStepMs() / args.bench_steps], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
        alg, active = sh.algorithmicBytes()
        out.update({"bench_size": d, "bench_blocks": M, "ms_per_step": ms, "cell_updates_per_s": d ** 3 / (ms * 1e-3),
                    "alg_GBps_all_ranks": alg / (ms * 1e-3) / 1e9, "steady": bool(sh.counters()[7]), "launches": int(sh.counters()[6])})
        dist.barrier()
        sh.close()
    if rank == 0:
        print(json.dumps(out), flush=True)
    dist.barrier()
    dist.destroy_process_group()
    sys.exit(0 if out["bit_exact_vs_single_gpu"] else 1)


if __name__ == "__main__":
    main()
