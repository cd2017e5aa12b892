"""Reader for the dump container written by oracle/ref_harness/harness.cu
([u32 name_len][name][u32 dtype][u64 count][payload])* and a runner for the harness."""
import json
import os
import struct
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF_DIR = os.path.join(ROOT, "oracle", "_ref")
_DT = {0: np.float32, 1: np.int32, 2: np.uint8, 3: np.uint64}


def read_dump(path):
    out = {}
    with open(path, "rb") as f:
        data = f.read()
    o = 0
    while o < len(data):
        (nl,) = struct.unpack_from("<I", data, o)
        o += 4
        name = data[o:o + nl].decode()
        o += nl
        dt, cnt = struct.unpack_from("<IQ", data, o)
        o += 12
        dtype = np.dtype(_DT[dt])
        out[name] = np.frombuffer(data, dtype=dtype, count=cnt, offset=o).copy()
        o += cnt * dtype.itemsize
    return out


def harness_path(fma):
    return os.path.join(REF_DIR, "ref_harness" if fma else "ref_harness_nofma")


def have_harness():
    return os.path.exists(harness_path(True)) and os.path.exists(harness_path(False))


def run_harness(fma=False, out=None, timeout=1800, **kw):
    """Runs the reference's own CUDA solver headless; returns (list of JSON lines, dump dict or None)."""
    cmd = [harness_path(fma)] + [f"{k}={v}" for k, v in kw.items()]
    if out:
        cmd.append(f"out={out}")
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=timeout)
    if r.returncode != 0:
        raise RuntimeError(f"ref_harness failed ({r.returncode}): {r.stderr[-2000:]}\n{r.stdout[-2000:]}")
    lines = [json.loads(l) for l in r.stdout.splitlines() if l.startswith("{")]
    return lines, (read_dump(out) if out else None), r.stdout
