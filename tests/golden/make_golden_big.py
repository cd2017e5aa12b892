"""Golden fixtures at the sizes bench.py measures (BASELINE.json configs[1], configs[2], the N-GPU bench scenes).

Run ON THE GPU BOX:

    gpurun -- 'python tests/golden/make_golden_big.py gpurun_out/golden_big'

Unlike make_golden.py the CPU oracle is NOT run here (1.7 s per 512^3 step): these fixtures pin the CUDA path
directly against the reference's own CUDA kernels (oracle/_ref/ref_harness_nofma: the reference sources built with
-fmad=false, its one racy kernel issued in rank order, see oracle/ref_harness/harness.cu).  Per case:

  * <case>.npz  SHA-256 of every canonical field and of the canonical block map (tests/_canon.py), SHA-256 and
                FNV-1a-64 of the RAW density / velocity arrays (slot order; the harness' own `final_digest`),
                a strided sample of each field, level loads and move limits
  * the "digest only" cases (the 4- and 8-GPU bench scenes) keep nothing but the harness' FNV-1a digest of the raw
    density + velocity arrays: their dumps would be 7-15 GB.

summary_big.json records, per case, the digest of the reference built with default flags (FMA contraction on) next
to the -fmad=false one, and — BASELINE.md's 100-step gate — rel-L2 / max-abs of default-flags vs -fmad=false fields.
"""
import hashlib
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from tests import _canon, _refio  # noqa: E402

# name, (gx, gy, gz), M, solids, steps, dump?
CASES = [
    ("d256_m65536", (256, 256, 256), 65536, 0, 100, True),                    # BASELINE configs[1]: 100 steps
    ("d512_m524288_solids", (512, 512, 512), 524288, 1, 140, True),           # configs[2] = bench.py's N=1 scene after its pre-roll
    ("d512x512x1024_m1048576_solids", (512, 512, 1024), 1048576, 1, 140, True),   # bench.py --gpus 2 scene
    ("d512x512x2048_m2097152_solids", (512, 512, 2048), 2097152, 1, 140, False),  # --gpus 4
    ("d512x512x4096_m4194304_solids", (512, 512, 4096), 4194304, 1, 140, False),  # --gpus 8
]
SAMPLE_STRIDE = 4099
FIELDS = ["density", "velocity", "fluidity", "pressure", "t_pressure", "divergence"]


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def cmp_f(a, b):
    a = np.asarray(a, dtype=np.float32).ravel().astype(np.float64)
    b = np.asarray(b, dtype=np.float32).ravel().astype(np.float64)
    d = np.abs(a - b)
    nb = float(np.sqrt(np.sum(b ** 2)))
    return {"max_abs": float(d.max()), "rel_l2": float(np.sqrt(np.sum(d ** 2)) / nb) if nb > 0 else 0.0, "max_ref": float(np.abs(b).max())}


def ref_topo(dump, pre):
    return dict(pos=dump[pre + "positions"].reshape(-1, 3), level=dump[pre + "levels"], parent=dump[pre + "parent"],
                child=dump[pre + "children"].reshape(-1, 8), apron=dump[pre + "apron"].reshape(-1, 216))


def main(outdir):
    os.makedirs(outdir, exist_ok=True)
    only = set(sys.argv[2:])
    spath = os.path.join(outdir, "summary_big.json")
    summary = {"cases": {}, "when": time.strftime("%Y-%m-%dT%H:%M:%SZ", time.gmtime())}
    for name, (gx, gy, gz), M, solids, steps, dump in CASES:
        if only and name not in only:
            continue
        t0 = time.time()
        tmp = f"/tmp/{name}"
        kw = dict(grid="dcgrid", gx=gx, gy=gy, gz=gz, M=M, solids=solids, steps=steps, schedule="project", serialize=1)
        lines_nf, ref, _ = _refio.run_harness(fma=False, out=(tmp + "_nofma.bin") if dump else None, timeout=3000, **kw)
        entry = {"gx": gx, "gy": gy, "gz": gz, "M": M, "solids": solids, "steps": steps,
                 "ref_nofma": {k: lines_nf[0][k] for k in ("ms_per_step", "advect_velocity_ms", "adapt_topology_ms", "project_ms",
                                                           "advect_density_ms", "final_digest")}}
        npz = {"fnv_raw_density_velocity": np.frombuffer(bytes.fromhex(lines_nf[0]["final_digest"]), dtype=np.uint8)}
        if dump:
            # the reference as its own CMake builds it (FMA contraction on): BASELINE.md's tolerance gate
            lines_f, ref_fma, _ = _refio.run_harness(fma=True, out=tmp + "_fma.bin", timeout=3000, **kw)
            entry["ref_fma_digest"] = lines_f[0]["final_digest"]
            entry["ref_fma_ms_per_step"] = lines_f[0]["ms_per_step"]
            same_map = all(np.array_equal(ref_fma["final/" + k], ref["final/" + k]) for k in ("positions", "levels", "parent", "children"))
            entry["ref_fma_same_block_map"] = bool(same_map)
            entry["ref_fma_vs_ref_nofma"] = {f: cmp_f(ref_fma["final/" + f], ref["final/" + f]) for f in FIELDS} if same_map else None
            del ref_fma
            os.remove(tmp + "_fma.bin")
            rt = ref_topo(ref, "final/")
            rf = {f: ref["final/" + f] for f in FIELDS}
            for f in ("density", "velocity"):
                npz["sha_raw_" + f] = np.frombuffer(bytes.fromhex(sha(np.ascontiguousarray(rf[f], dtype=np.float32))), dtype=np.uint8)
            ca = _canon.canonical(rt, rf)
            for k in ("blocks", "parent", "child", "apron"):
                npz["sha_topo_" + k] = np.frombuffer(bytes.fromhex(sha(ca[k].astype(np.int32))), dtype=np.uint8)
            for f in FIELDS:
                arr = np.ascontiguousarray(ca[f], dtype=np.float32)
                npz["sample_" + f] = arr.reshape(-1)[::SAMPLE_STRIDE].copy()
                npz["sha_" + f] = np.frombuffer(bytes.fromhex(sha(arr)), dtype=np.uint8)
            npz["loads"] = ref["final/block_loads"]
            npz["move_limit"] = ref["final/move_limit"]
            entry["loads"] = ref["final/block_loads"].tolist()
            entry["active_blocks"] = int(np.count_nonzero(rt["level"] != 0xFF))
            del ref, rt, rf, ca
            os.remove(tmp + "_nofma.bin")
        npz["meta"] = np.frombuffer(json.dumps({"grid": "dcgrid", "gx": gx, "gy": gy, "gz": gz, "d": gx, "M": M, "solids": solids, "steps": steps,
                                                "schedule": "project", "sample_stride": SAMPLE_STRIDE, "big": True, "digest_only": not dump,
                                                "source": "oracle/_ref/ref_harness_nofma serialize=1 (reference CUDA, -fmad=false) on B200"}).encode(),
                                    dtype=np.uint8)
        np.savez_compressed(os.path.join(outdir, name + ".npz"), **npz)
        entry["seconds"] = round(time.time() - t0, 1)
        summary["cases"][name] = entry
        print(name, json.dumps(entry)[:1200], flush=True)
        with open(spath, "w") as f:
            json.dump(summary, f, indent=1)


if __name__ == "__main__":
    main(sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "gpurun_out", "golden_big"))
