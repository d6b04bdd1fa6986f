#!/usr/bin/env python3
"""Same B200, same rays, three implementations of the EFIT RK4 step (run on the GPU box):
  1. the reference's own gpu::cuda_context (integration/_build/ref_driver_cuda)
  2. the reference front end on the drop-in b200_context (integration/_build/ref_driver_b200)
  3. the full B200-native back end (bench.py numbers are printed separately)
plus the reference CPU path.  Prints one JSON line per arm."""
import json
import os
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np                                  # noqa: E402
from graph_framework_b200 import workloads          # noqa: E402

disp = sys.argv[1] if len(sys.argv) > 1 else "extra_ordinary_wave"
n = int(sys.argv[2]) if len(sys.argv) > 2 else 100000
steps = int(sys.argv[3]) if len(sys.argv) > 3 else 100
state = workloads.efit_ensemble(n, seed=0)
env = dict(os.environ, GFB_EFIT_FILE=os.path.join(ROOT, "tests", "golden", "efit.gfbt"))
with tempfile.TemporaryDirectory() as d:
    fin = os.path.join(d, "in.bin")
    workloads.pack(state).tofile(fin)
    for arm, exe, threads in (("reference cuda_context", "integration/_build/ref_driver_cuda", 1),
                              ("reference front end + b200_context", "integration/_build/ref_driver_b200", 1),
                              ("reference cpu path", "oracle/_ref/ref_driver", os.cpu_count())):
        path = os.path.join(ROOT, exe)
        if not os.path.exists(path):
            print(json.dumps({"arm": arm, "unavailable": exe}))
            continue
        try:
            out = subprocess.run([path, "bench", disp, "efit", str(n), "2e-5", str(steps), str(threads), fin],
                                 cwd=ROOT, env=env, capture_output=True, text=True, timeout=1500)
            line = [l for l in out.stdout.splitlines() if l.startswith("{")]
            if out.returncode != 0 or not line:
                print(json.dumps({"arm": arm, "failed": out.returncode, "stderr": out.stderr[-400:], "stdout": out.stdout[-400:]}))
                continue
            r = json.loads(line[-1])
            r["arm"] = arm
            r["dispersion"] = disp
            print(json.dumps(r))
        except subprocess.TimeoutExpired:
            print(json.dumps({"arm": arm, "failed": "timeout"}))
