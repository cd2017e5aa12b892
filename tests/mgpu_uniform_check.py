"""Multi-process check of the z-slab sharded uniform solver over NVLink peer mappings.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
        tests/mgpu_uniform_check.py [--size 128] [--steps 8] [--bench-size 512] [--bench-steps 10]

Every rank runs (a) its slab of the N-rank decomposition (one process per GPU, CUDA-IPC peer mappings,
flag barriers) and (b) the plain single-GPU solver on the whole grid, and compares its slab bit for bit.
Then the sharded solver is timed on a larger grid (device time, max over ranks).  Not collected by pytest:
needs >= 2 GPUs (gpurun --gpus 2)."""
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
    ap.add_argument("--steps", type=int, default=8)
    ap.add_argument("--bench-size", type=int, default=512)
    ap.add_argument("--bench-steps", type=int, default=10)
    args = ap.parse_args()
    import torch
    import torch.distributed as dist

    from dcgrid_b200 import FluidSimulationUniform, FluidSimulationUniformSharded, scene_params

    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    ok = True
    for d, solids, schedule in ((args.size, True, "project"), (64, False, "local")):
        p = scene_params(d, solids=solids)
        sh = FluidSimulationUniformSharded((d, d, d), p, world, rank=rank, nlocal=1, device=local, dist=dist)
        one = FluidSimulationUniform((d, d, d), p, device=local)
        for _ in range(args.steps):
            for sim in (sh, one):
                sim.advectVelocity(); sim.adaptTopology()
                sim.project() if schedule == "project" else sim.projectLocal()
                sim.advectDensity()
        z0, z1 = sh.z_range
        n = d * d
        for f in ("density", "velocity", "pressure", "divergence"):
            a = sh.field(f)
            b = one.field(f)[z0 * n:z1 * n]
            same = np.array_equal(np.ascontiguousarray(a).view(np.uint32), np.ascontiguousarray(b).view(np.uint32))
            ok = ok and same
            if not same:
                print(f"[rank {rank}] MISMATCH {f} d={d}: max abs {np.abs(a - b).max()}", flush=True)
        tot_s, tot_1 = sh.totalDensity(), one.totalDensity()
        ok = ok and abs(tot_s - tot_1) <= 1e-9 * max(1.0, abs(tot_1))
        dist.barrier()
        sh.close(); one.close()
    flag = torch.tensor([1 if ok else 0], device="cuda")
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    # ---- timing of the sharded solver on a larger grid (--bench-size 0: parity only)
    out = {"check": "mgpu_uniform", "n_gpus": world, "bit_exact_vs_single_gpu": bool(flag.item()), "parity_size": args.size}
    if args.bench_size > 0:
        d = args.bench_size
        p = scene_params(d, solids=True)
        sh = FluidSimulationUniformSharded((d, d, d), p, world, rank=rank, nlocal=1, device=local, dist=dist)
        sh.step(3)
        dist.barrier(); torch.cuda.synchronize()
        sh.step(args.bench_steps)
        ms = torch.tensor([sh.lastStepMs()], dtype=torch.float64, device="cuda")
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        bytes_local, _ = sh.algorithmicBytes()
        bt = torch.tensor([bytes_local], dtype=torch.float64, device="cuda")
        dist.all_reduce(bt)
        ctr = sh.counters()
        per = float(ms.item()) / args.bench_steps
        out.update({"bench_size": d, "ms_per_step": per, "cell_updates_per_s": d ** 3 / (per * 1e-3),
                    "alg_GBps_all_ranks": float(bt.item()) / (per * 1e-3) / 1e9, "barriers": int(ctr[7]), "launches": int(ctr[6])})
        sh.close()
    if rank == 0:
        print(json.dumps(out), flush=True)
    dist.barrier()
    dist.destroy_process_group()
    sys.exit(0 if flag.item() else 1)


if __name__ == "__main__":
    main()
