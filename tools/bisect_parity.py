#!/usr/bin/env python
"""Locates the first step (and field / cells) where the CUDA path and the reference's own CUDA kernels
(oracle/_ref/ref_harness_nofma serialize=1) stop being bit-identical.  Run on the GPU box:

    python tools/bisect_parity.py --d 512 --M 524288 --solids 1 --steps 140 [--variants]

1. per-step FNV digests of raw density + velocity from both sides (harness trace=1);
2. at the first mismatching step: full dumps of both, compared array by array (block pool, then fields),
   mismatches broken down by level;
3. --variants: the digest after `steps` steps of a few dcg_options variants (which code path matters).
"""
import argparse
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from dcgrid_b200 import FluidSimulationDCGrid, fnv1a64, scene_params  # noqa: E402
from tests import _refio  # noqa: E402


def bits(a):
    return np.ascontiguousarray(a, dtype=np.float32).view(np.uint32)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--d", type=int, default=512)
    ap.add_argument("--gz", type=int, default=0)
    ap.add_argument("--M", type=int, default=524288)
    ap.add_argument("--solids", type=int, default=1)
    ap.add_argument("--steps", type=int, default=140)
    ap.add_argument("--variants", action="store_true")
    ap.add_argument("--opt", action="append", default=[])
    a = ap.parse_args()
    size = (a.d, a.d, a.gz or a.d)
    opts = {k: int(v) for k, v in (o.split("=") for o in a.opt)}
    kw = dict(grid="dcgrid", gx=size[0], gy=size[1], gz=size[2], M=a.M, solids=a.solids, schedule="project", serialize=1)
    lines, _, _ = _refio.run_harness(fma=False, steps=a.steps, trace=1, timeout=3000, **kw)
    trace = [int(t, 16) for t in lines[0]["trace"]]
    print("reference trace:", len(trace), "steps; final", lines[0]["final_digest"], flush=True)
    p = scene_params(*size, solids=bool(a.solids))
    sim = FluidSimulationDCGrid(size, a.M, p, options=opts or None)
    first_bad = None
    for s in range(a.steps):
        sim.step(1)
        got = fnv1a64(sim.field("density"), sim.field("velocity"))
        ok = got == trace[s]
        if s % 10 == 0 or not ok:
            print(f"step {s}: ours {got:016x} ref {trace[s]:016x} {'ok' if ok else 'MISMATCH'} counters {sim.counters().tolist()}", flush=True)
        if not ok:
            first_bad = s
            break
    del sim
    if first_bad is None:
        print(json.dumps({"result": "bit-identical", "steps": a.steps}))
    else:
        n = first_bad + 1
        out = f"/tmp/bisect_{n}.bin"
        _, ref, _ = _refio.run_harness(fma=False, steps=n, out=out, timeout=3000, **kw)
        os.remove(out)
        sim = FluidSimulationDCGrid(size, a.M, p, options=opts or None)
        if n > 1:
            sim.step(n - 1)
        q_before = sim.field("density").copy()
        v_before = sim.field("velocity").copy()
        sim.advectVelocity()
        v_adv = sim.field("velocity").copy()
        sim.adaptTopology()
        v_adapt = sim.field("velocity").copy()
        sim.project()
        mine = {f: sim.field(f).copy() for f in ("pressure", "t_pressure", "divergence")}
        mine["velocity_projected"] = sim.field("velocity").copy()
        sim.advectDensity()
        for f in ("density", "velocity", "fluidity"):
            mine[f] = sim.field(f).copy()
        topo = sim.topology()
        lv = ref["final/levels"]
        print("block pool:")
        for k, rk, shape in (("level", "levels", None), ("pos", "positions", (-1, 3)), ("parent", "parent", None), ("child", "children", (-1, 8)),
                             ("apron", "apron", (-1, 216))):
            r = ref["final/" + rk]
            if shape:
                r = r.reshape(shape)
            m = topo[k].astype(np.uint64) != r.astype(np.uint64) if k != "pos" else topo[k] != r
            if k in ("pos",):
                m = m & (lv != 0xFF)[:, None]
            print(f"  {k}: {int(np.count_nonzero(m))} mismatching entries", flush=True)
        print("loads ours", sim.levelTable()["loads"].tolist(), "ref", ref["final/block_loads"].tolist())
        lvc = np.repeat(lv, 64)
        pos = ref["final/positions"].reshape(-1, 3)
        for f in ("divergence", "pressure", "t_pressure", "density", "velocity", "fluidity"):
            r = ref["final/" + f]
            g = mine[f]
            comps = 3 if f == "velocity" else 1
            m = (bits(g).reshape(-1, comps) != bits(r).reshape(-1, comps)).any(axis=1)
            print(f"field {f}: {int(np.count_nonzero(m))} mismatching cells of {m.size}")
            if m.any():
                for l in range(int(lv[lv != 0xFF].max()) + 1):
                    c = int(np.count_nonzero(m & (lvc == l)))
                    if c:
                        print(f"    level {l}: {c}")
                print("    free-slot cells:", int(np.count_nonzero(m & (lvc == 0xFF))))
                idx = np.flatnonzero(m)[:12]
                for i in idx:
                    b, c = divmod(int(i), 64)
                    print(f"    cell {i} slot {b} level {lv[b]} pos {pos[b].tolist()} bits {c:06b} ours {g.reshape(-1, comps)[i]} ref {r.reshape(-1, comps)[i]}")
        # stage attribution on our side: which of our own stages already differs is unknowable without reference dumps of
        # the intermediate state; print the magnitude of what the step did for orientation
        print("ours: |v_adv - v_before| max", float(np.abs(v_adv - v_before).max()), "|v_adapt - v_adv| max", float(np.abs(v_adapt - v_adv).max()))
        del q_before
    if a.variants:
        for o in ({}, {"no_resort": 1}, {"host_selection": 1}, {"advect_no_fuse": 1}, {"no_pdl": 1}, {"jacobi": 1}, {"stencil": 1}, {"advect": 1},
                  {"coarse_in_gmem": 1, "zero_all": 1}, {"resort_every": -1}):
            sim = FluidSimulationDCGrid(size, a.M, p, options=o or None)
            sim.step(a.steps)
            print("variant", o, f"{fnv1a64(sim.field('density'), sim.field('velocity')):016x}", "ref", lines[0]["final_digest"], flush=True)
            del sim


if __name__ == "__main__":
    main()
