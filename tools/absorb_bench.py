"""What does device-resident absorption cost?  O-mode EFIT, N rays, blocks of 100 RK4 steps:
stepping only / stepping + records / stepping + weak damping + power + deposition (+ records).
Timed with CUDA events on the launching stream; wall clock for the variants with host copies.
usage (GPU box): python tools/absorb_bench.py [rays] [blocks]"""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from graph_framework_b200.rays import RayTracer, pinned_empty
from graph_framework_b200 import workloads

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1000000
blocks = int(sys.argv[2]) if len(sys.argv) > 2 else 20
sub = 100
out = {"rays": n, "blocks": blocks, "sub_steps": sub, "dispersion": "ordinary_wave", "equilibrium": "efit", "dt": 1.0e-3}
for absorb in (0, 1):
    tr = RayTracer("ordinary_wave", "efit", n, 1.0e-3, options="absorption=%d" % absorb)
    tr.set_state(workloads.efit_ensemble(n, seed=0))
    tr.init("kx"); tr.compile()
    tr.step(sub); tr.wait()
    if not absorb:
        tr.timer_start(); tr.step(sub*blocks); ms = tr.timer_stop()
        out["step_only_ms_per_block"] = ms/blocks
        out["step_only_ray_steps_per_s"] = n*sub*blocks/(ms*1e-3)
        rec = pinned_empty((blocks, 9, n))
        tr.trace(2, sub, out=rec[:2])
        t0 = time.perf_counter(); tr.trace(blocks, sub, out=rec); t1 = time.perf_counter()
        out["trace_records_ms_per_block"] = 1e3*(t1 - t0)/blocks
        out["trace_records_ray_steps_per_s"] = n*sub*blocks/(t1 - t0)
    else:
        lo, hi, bins = (1.0, -1.0, -1.0), (2.6, 1.0, 1.0), (64, 16, 16)
        rec, ab = pinned_empty((blocks, 9, n)), pinned_empty((blocks, 3, n))
        tr.trace_absorb(2, sub, bins=bins, lo=lo, hi=hi, records=False, absorbed_out=ab[:2])
        l0 = tr.launch_count()
        t0 = time.perf_counter(); _, a, prof = tr.trace_absorb(blocks, sub, bins=bins, lo=lo, hi=hi, records=False, absorbed_out=ab); t1 = time.perf_counter()
        out["absorb_launches_per_block"] = (tr.launch_count() - l0)/blocks
        out["trace_absorb_ms_per_block"] = 1e3*(t1 - t0)/blocks
        out["trace_absorb_ray_steps_per_s"] = n*sub*blocks/(t1 - t0)
        out["median_transmitted_power"] = float(np.nanmedian(a[-1, 1]))
        out["finite_power_fraction"] = float(np.isfinite(a[-1, 1]).mean())
        t0 = time.perf_counter(); tr.trace_absorb(blocks, sub, bins=bins, lo=lo, hi=hi, records=True, records_out=rec, absorbed_out=ab); t1 = time.perf_counter()
        out["trace_absorb_records_ms_per_block"] = 1e3*(t1 - t0)/blocks
        tr.timer_start()
        tr.trace_absorb(blocks, 0, bins=bins, lo=lo, hi=hi, records=False, absorbed_out=ab)   # sub_steps = 0: absorption kernels only
        out["absorption_kernels_ms_per_record"] = tr.timer_stop()/blocks
    tr.close()
print(json.dumps(out))
