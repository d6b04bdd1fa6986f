#!/usr/bin/env python3
"""Convert the reference's netCDF-4 fixtures to GFBT (run in the build container only).

TEST INFRASTRUCTURE.  Reads /root/reference/graph_tests/{efit,efit_gold,vmec}.nc
with the package's HDF5 reader and writes tests/golden/*.gfbt, which are
committed so the GPU box (no /root/reference there) has the same tables.
"""
import os
import sys
HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)
from graph_framework_b200.tools.gfbt import nc_to_gfbt  # noqa: E402

REF = os.environ.get("GFB_REFERENCE", "/root/reference")
for stem in ("efit", "efit_gold", "vmec"):
    src = os.path.join(REF, "graph_tests", stem + ".nc")
    dst = os.path.join(ROOT, "tests", "golden", stem + ".gfbt")
    arrs = nc_to_gfbt(src, dst)
    print("%s -> %s (%d variables, %d bytes)" % (src, dst, len(arrs), os.path.getsize(dst)))
