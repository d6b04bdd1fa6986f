"""Where does the xrays-style trace spend its time at 1e5 rays? (GPU box)"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from graph_framework_b200.rays import RayTracer
from graph_framework_b200 import workloads

n = int(sys.argv[1]) if len(sys.argv) > 1 else 100000
for opts in ("", "block=64", "block=256"):
    tr = RayTracer("ordinary_wave", "efit", n, 2.0e-5, options=opts)
    tr.set_state(workloads.efit_ensemble(n, seed=0))
    tr.init("kx"); tr.compile()
    print(opts or "default", tr.kernel_stats())
    for blk in (100, 1000):
        tr.step(blk); tr.wait()
        t0 = time.perf_counter(); tr.step(blk*10); tr.wait(); t1 = time.perf_counter()
        print("  step(%d) x10: %.3e ray-steps/s" % (blk, n*blk*10/(t1 - t0)))
    for label in ("first", "second"):
        t0 = time.perf_counter(); rec = tr.trace(10, 1000); t1 = time.perf_counter()
        print("  %s trace(10,1000): %.3e ray-steps/s; finite %s" % (label, n*1e4/(t1 - t0), np.isfinite(rec[-1]).mean()))
    t0 = time.perf_counter(); rec = tr.trace(10, 1000, out=rec); t1 = time.perf_counter()
    print("  trace into the same pinned buffer: %.3e ray-steps/s" % (n*1e4/(t1 - t0)))
    s = tr.get_state()
    print("  R range after %g s: %.3f..%.3f" % (s["t"][0], np.hypot(s["x"], s["y"]).min(), np.hypot(s["x"], s["y"]).max()))
    tr.close()
