"""CPU checks of bench.py's bookkeeping: the digest that ties committed ncu figures to the kernel sources they were taken on."""
import json
import os
import shutil

import bench

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_translation_unit_digests_are_independent(tmp_path, monkeypatch):
    for d in ("dcgrid_b200/csrc", "include"):
        shutil.copytree(os.path.join(ROOT, d), tmp_path / d)
    monkeypatch.setattr(bench, "ROOT", str(tmp_path))
    whole, dc, un = bench.csrc_digest(), bench.csrc_digest("dcgrid"), bench.csrc_digest("uniform")
    assert len({whole, dc, un}) == 3
    with open(tmp_path / "dcgrid_b200/csrc/uniform.cu", "a") as f:
        f.write("// touched\n")
    assert bench.csrc_digest("dcgrid") == dc, "uniform.cu is not part of the DCGrid translation unit"
    assert bench.csrc_digest("uniform") != un and bench.csrc_digest() != whole
    with open(tmp_path / "dcgrid_b200/csrc/common.cuh", "a") as f:
        f.write("// touched\n")
    assert bench.csrc_digest("dcgrid") != dc, "a shared header belongs to both translation units"


def test_committed_traffic_file_names_its_sources():
    with open(os.path.join(ROOT, "profiles", "dominant_kernel_traffic.json")) as f:
        tj = json.load(f)
    assert tj["dram_bytes_per_launch"] > 0 and len(tj["tu_digest"]) == 16 and tj["kernel"] == "k_dc_jacobi_pipe8"
    assert set(tj["all_hot_kernels"]) >= {"k_dc_jacobi_pipe8", "k_dc_divergence_pipe"}
