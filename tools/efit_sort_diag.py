"""Does sorting rays by EFIT (R, Z) cell help the EFIT ray kernels?  (GPU box)"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from graph_framework_b200.rays import RayTracer
from graph_framework_b200 import workloads
n = 1000000
for disp in ("extra_ordinary_wave", "cold_plasma"):
    base = workloads.efit_ensemble(n, seed=0)
    base = {k: np.ascontiguousarray(np.broadcast_to(np.asarray(v, dtype=np.float64), (n,))) for k, v in base.items()}
    for mode in ("caller order", "sorted by cell", "device binning"):
        s = base
        if mode == "sorted by cell":
            r = np.hypot(base["x"], base["y"])
            ir = np.clip((r - 0.84)/0.0265625, 0, 63).astype(np.int64)
            iz = np.clip((base["z"] + 1.6)/0.05, 0, 63).astype(np.int64)
            order = np.argsort(ir*64 + iz, kind="stable")
            s = {k: v[order] for k, v in base.items()}
        tr = RayTracer(disp, "efit", n, 2.0e-5, options=None if mode == "device binning" else "bin_rays=0")
        tr.set_state(s); tr.init("kx"); tr.compile()
        tr.step(100); tr.wait()
        ms = 0.0
        for _ in range(5):
            tr.timer_start(); tr.step(100); ms += tr.timer_stop()
        print(disp, mode, "%.3e ray-steps/s" % (n*500/(ms*1e-3)), "%.2f ms" % (ms/5))
        tr.close()
