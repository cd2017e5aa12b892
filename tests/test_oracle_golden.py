"""CPU-only: the oracle (oracle/liboracle.so) against the golden vectors in tests/golden/.

The vectors are outputs of the REFERENCE's own CUDA kernels (oracle/_ref/ref_harness_nofma, i.e. the
reference sources compiled with -fmad=false) captured on a B200 by tests/golden/make_golden.py;
DCGrid cases were run with the reference's one racy kernel serialised in rank order (see
oracle/ref_harness/harness.cu).  Bar: bit-exact — full arrays for the tiny cases, SHA-256 digests for
the larger ones."""
import hashlib
import json
import os

import numpy as np
import pytest

from dcgrid_b200.params import scene_params
from tests import _canon
from tests._oracle import Oracle

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
FIELDS = ["density", "velocity", "fluidity", "pressure", "t_pressure", "divergence"]
CASES = sorted(f[:-4] for f in os.listdir(GOLDEN) if f.endswith(".npz"))
# cases that take long on few CPU cores are still run, but with the cheapest first
FAST = [c for c in CASES if c.startswith(("u32", "d32", "d64"))]
SLOW = [c for c in CASES if c not in FAST]


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def run_oracle(meta):
    grid, d, M, solids, steps, schedule = (meta[k] for k in ("grid", "d", "M", "solids", "steps", "schedule"))
    p = scene_params(d, solids=bool(solids))
    o = Oracle(p, M if grid == "dcgrid" else 0)
    if schedule.startswith("jacobi"):
        o.set_jacobi_schedule(2, 1, int(schedule[6:]))
    out = {}
    for s in range(steps):
        o.advect_velocity()
        o.adapt_topology()
        o.project() if schedule == "project" else o.project_local()
        if s == steps - 1:
            for f in ("pressure", "t_pressure", "divergence"):
                out[f] = o.field(f).copy()
        o.advect_density()
    for f in ("density", "velocity", "fluidity"):
        out[f] = o.field(f).copy()
    if grid == "dcgrid":
        out["topo"] = o.topology()
        out["loads"] = o.level_table()["loads"].copy()
        out["move_limit"] = o.move_limits()
    return out


@pytest.mark.parametrize("case", FAST + SLOW)
def test_oracle_reproduces_reference_cuda_outputs(case):
    z = np.load(os.path.join(GOLDEN, case + ".npz"))
    meta = json.loads(bytes(z["meta"]).decode())
    out = run_oracle(meta)
    if meta["grid"] == "dcgrid":
        canon = _canon.canonical(out["topo"], {f: out[f] for f in FIELDS})
        np.testing.assert_array_equal(out["loads"], z["loads"])
        np.testing.assert_array_equal(out["move_limit"], z["move_limit"])
        for k in ("blocks", "parent", "child", "apron"):
            arr = canon[k].astype(np.int32)
            assert sha(arr) == bytes(z["sha_topo_" + k]).hex(), f"topology {k}"
            if "topo_" + k in z:
                np.testing.assert_array_equal(arr, z["topo_" + k])
        got = {f: np.ascontiguousarray(canon[f], dtype=np.float32) for f in FIELDS}
    else:
        got = {f: np.ascontiguousarray(out[f], dtype=np.float32) for f in FIELDS}
    for f in FIELDS:
        if f in z:
            np.testing.assert_array_equal(got[f].view(np.uint32).ravel(), z[f].view(np.uint32).ravel(), err_msg=f)
        else:
            np.testing.assert_array_equal(got[f].reshape(-1)[::meta["sample_stride"]].view(np.uint32),
                                          z["sample_" + f].view(np.uint32), err_msg=f"{f} (strided sample)")
        assert sha(got[f]) == bytes(z["sha_" + f]).hex(), f"{f}: SHA-256 of the full array"


def test_pinning_summary_says_bit_exact():
    """The committed summary of the pinning run: oracle == reference CUDA (-fmad=false) in every field of every case."""
    s = json.load(open(os.path.join(GOLDEN, "summary.json")))
    assert len(s["cases"]) >= 9
    for name, c in s["cases"].items():
        for f, r in c["oracle_vs_ref_nofma"].items():
            assert r["bit_mismatches"] == 0, (name, f)
        if c["grid"] == "dcgrid":
            assert all(v == 0 for v in c["topology_raw_mismatches"].values()), name
            assert len(set(c["ref_serialized_rep_digests"])) == 1, name
