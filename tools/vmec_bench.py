#!/usr/bin/env python3
"""VMEC O-mode / cold-plasma RK4 throughput on one GPU (BASELINE configs[3] per-GPU share)."""
import json
import sys
import os
import time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from graph_framework_b200 import workloads                 # noqa: E402
from graph_framework_b200.rays import RayTracer            # noqa: E402

disp = sys.argv[1] if len(sys.argv) > 1 else "cold_plasma"
n = int(sys.argv[2]) if len(sys.argv) > 2 else 1250000       # 10^7 rays / 8 GPUs
steps = int(sys.argv[3]) if len(sys.argv) > 3 else 10
opts = sys.argv[4] if len(sys.argv) > 4 else ""
state = workloads.vmec_states(n, seed=0)
if os.environ.get("GFB_SORT_RAYS"):
    order = __import__("numpy").argsort(state["x"], kind="stable")      # radial cell locality experiment
    state = {k: v[order] for k, v in state.items()}
t0 = time.perf_counter()
if not os.environ.get("GFB_BIN_RAYS"):
    opts += " bin_rays=0"                                                # default: binned by radial cell, re-sorted every 50 steps
elif os.environ.get("GFB_REBIN"):
    opts += " bin_rays=" + os.environ["GFB_REBIN"]
tr = RayTracer(disp, "vmec", n, 1.0e-4, options=("fused_steps=%d " % steps) + opts)
tr.set_state(state)
t1 = time.perf_counter()
tr.init("kx")
t2 = time.perf_counter()
tr.compile()
t3 = time.perf_counter()
tr.step(steps)
tr.wait()
ms = 0.0
reps = 3
for _ in range(reps):
    tr.timer_start()
    tr.step(steps)
    ms += tr.timer_stop()
st = tr.get_state()
stats = tr.kernel_stats()
import numpy as np
print(json.dumps({"workload": "vmec %s RK4" % disp, "rays": n, "steps_per_launch": steps,
                  "ray_steps_per_s": n*steps*reps/(ms*1e-3), "ms_per_launch": ms/reps,
                  "finite": bool(np.isfinite(st["x"]).all()), "max_residual": float(np.max(st["residual"])),
                  "binned": bool(os.environ.get("GFB_BIN_RAYS")), "kernel": stats, "setup_s": t1 - t0, "newton_s": t2 - t1, "jit_s": t3 - t2, "options": opts}))
