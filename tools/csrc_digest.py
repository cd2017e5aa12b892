#!/usr/bin/env python
"""Prints the digest bench.py ties ncu figures to (SHA-256 over dcgrid_b200/csrc + include)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402

print(bench.csrc_digest(sys.argv[1] if len(sys.argv) > 1 else None))
