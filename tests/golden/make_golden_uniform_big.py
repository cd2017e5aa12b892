"""Golden digests of the DENSE solver at sizes that engage its z-marching kernels (x >= 128 cells) and, sharded, the slab
decomposition of BASELINE configs[4] — from the reference's own CUDA kernels (oracle/_ref/ref_harness_nofma, the reference
sources built with -fmad=false).  Run ON THE GPU BOX:

    gpurun -- 'python tests/golden/make_golden_uniform_big.py gpurun_out/golden_bigu'

Per case <case>.npz: FNV-1a-64 of the raw density + velocity arrays (the harness' `final_digest`), SHA-256 and a strided
sample of density / velocity / pressure / divergence for the cases that are dumped.  tests/test_golden_uniform_big_gpu.py
compares the CUDA path (one GPU, and 2 / 4 ranks of the slab decomposition in one process) bit for bit."""
import hashlib
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from tests import _refio  # noqa: E402

# name, (gx, gy, gz), solids, steps, schedule, dump?
CASES = [
    ("u256_solids_project", (256, 256, 256), 1, 40, "project", True),
    ("u256x128x512_local", (256, 128, 512), 0, 30, "local", True),
    ("u512_solids_project", (512, 512, 512), 1, 20, "project", False),
]
SAMPLE_STRIDE = 4099
FIELDS = ["density", "velocity", "pressure", "divergence"]


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def main(outdir):
    os.makedirs(outdir, exist_ok=True)
    summary = {"cases": {}, "when": time.strftime("%Y-%m-%dT%H:%M:%SZ", time.gmtime())}
    for name, (gx, gy, gz), solids, steps, schedule, dump in CASES:
        t0 = time.time()
        tmp = f"/tmp/{name}.bin"
        kw = dict(grid="uniform", gx=gx, gy=gy, gz=gz, solids=solids, steps=steps, schedule=schedule)
        lines, ref, _ = _refio.run_harness(fma=False, out=tmp if dump else None, timeout=3000, **kw)
        npz = {"fnv_raw_density_velocity": np.frombuffer(bytes.fromhex(lines[0]["final_digest"]), dtype=np.uint8)}
        if dump:
            for f in FIELDS:
                arr = np.ascontiguousarray(ref["final/" + f], dtype=np.float32)
                npz["sample_" + f] = arr.reshape(-1)[::SAMPLE_STRIDE].copy()
                npz["sha_" + f] = np.frombuffer(bytes.fromhex(sha(arr)), dtype=np.uint8)
            del ref
            os.remove(tmp)
        npz["meta"] = np.frombuffer(json.dumps({"grid": "uniform", "gx": gx, "gy": gy, "gz": gz, "solids": solids, "steps": steps, "schedule": schedule,
                                                "sample_stride": SAMPLE_STRIDE, "digest_only": not dump,
                                                "source": "oracle/_ref/ref_harness_nofma (reference CUDA, -fmad=false) on B200"}).encode(), dtype=np.uint8)
        np.savez_compressed(os.path.join(outdir, name + ".npz"), **npz)
        summary["cases"][name] = {"digest": lines[0]["final_digest"], "ref_ms_per_step": lines[0]["ms_per_step"], "seconds": round(time.time() - t0, 1)}
        print(name, summary["cases"][name], flush=True)
    with open(os.path.join(outdir, "summary_bigu.json"), "w") as f:
        json.dump(summary, f, indent=1)


if __name__ == "__main__":
    main(sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "gpurun_out", "golden_bigu"))
