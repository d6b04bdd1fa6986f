"""Does sorting particles by EFIT (R, Z) cell speed up the Boris push?  (GPU box)"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from graph_framework_b200.rays import BorisPusher
from graph_framework_b200 import workloads

n = int(sys.argv[1]) if len(sys.argv) > 1 else 20000000
p = dict(zip(("x", "y", "z", "ux", "uy", "uz"), workloads.boris_ensemble(n, seed=0)))
for mode in ("unsorted", "sorted", "device binning 100", "device binning 300", "device binning 1000"):
    s = p
    if mode == "sorted":
        r = np.hypot(p["x"], p["y"])
        ir = np.clip((r - 0.84)/0.0265625, 0, 63).astype(np.int64)
        iz = np.clip((p["z"] + 1.6)/0.05, 0, 63).astype(np.int64)
        order = np.argsort(ir*64 + iz, kind="stable")
        s = {k: v[order] for k, v in p.items()}
    push = BorisPusher("efit", n, dt=0.5, options="fused_steps=100 bin_rays=0")
    push.set_state(s["x"], s["y"], s["z"], s["ux"], s["uy"], s["uz"])
    push.compile()
    if mode.startswith("device binning"):
        push.set_binning((0.84, 0.84 + 64*0.0265625, 64), (-1.6, 1.6, 64), rebin_every=int(mode.split()[-1]))
    push.step(100)
    ms = 0.0
    reps = 12
    for _ in range(reps):
        push.timer_start(); push.step(100); ms += push.timer_stop()
    print(mode, "%.3e particle-steps/s" % (n*100*reps/(ms*1e-3)), "%.1f ms/launch" % (ms/reps))
    push.close() if hasattr(push, "close") else None
