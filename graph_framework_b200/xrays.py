"""xrays: the ray tracing stage of the reference's driver on the B200 back end.

    python -m graph_framework_b200.xrays --num_rays=100000 --num_times=100000 --sub_steps=100 \\
        --endtime=2.0 --dispersion=ordinary_wave --equilibrium=efit --equilibrium_file=efit.gfbt \\
        --init_kx --init_kx_mean=-700 --init_w_dist=normal --init_w_mean=700 --init_w_sigma=10 ... --use_cyl_xy

Option names and meaning follow /root/reference/graph_driver/xrays.cpp:955-1037 (trace_ray stage,
xrays.cpp:413-529): one host thread per device, rays split batch/extra (xrays.cpp:423-432), initial
conditions drawn per shard in the order w, kx, ky, kz, z, (x, y) (xrays.cpp:448-453), `--init_kx`
etc. select the component solved from the dispersion relation, every `sub_steps` steps a record is
written.  Differences: the random stream is numpy's (seeded with the shard index; the reference
uses std::mt19937_64), equilibrium files are GFBT (tools/gfbt.py converts netCDF), results are
result<shard>.gfbt with one (time, num_rays) variable per quantity, and the absorption and
power stages (complex arithmetic) are not part of this back end.
"""
import argparse
import threading
import time

import numpy as np

from ._lib import lib
from .rays import RayTracer, STATE, shard_offsets
from .tools.gfbt import write_trajectory

VARS = ("w", "kx", "ky", "kz", "x", "y", "z")


def parser():
    p = argparse.ArgumentParser(prog="xrays", description=__doc__, formatter_class=argparse.RawDescriptionHelpFormatter)
    p.add_argument("--verbose", action="store_true")
    p.add_argument("--num_times", type=int, default=1000)
    p.add_argument("--sub_steps", type=int, default=100)
    p.add_argument("--num_rays", type=int, default=1)
    p.add_argument("--endtime", type=float, default=1.0)
    p.add_argument("--solver", default="rk4", choices=["rk2", "rk4", "split_simplextic"])
    p.add_argument("--dispersion", default="ordinary_wave",
                   choices=["simple", "bohm_gross", "ordinary_wave", "extra_ordinary_wave", "cold_plasma"])
    p.add_argument("--equilibrium", default="efit", choices=["efit", "vmec", "slab", "slab_density", "slab_field",
                                                             "no_magnetic_field", "gaussian_density"])
    p.add_argument("--equilibrium_file", default=None)
    for v in VARS:
        p.add_argument("--init_%s_dist" % v, default="uniform", choices=["uniform", "normal"])
        p.add_argument("--init_%s_mean" % v, type=float, default=0.0)
        p.add_argument("--init_%s_sigma" % v, type=float, default=0.0)
    for v in ("kx", "ky", "kz"):
        p.add_argument("--init_%s" % v, action="store_true", help="solve this component from the dispersion relation")
    p.add_argument("--use_cyl_xy", action="store_true")
    p.add_argument("--seed", action="store_true", help="fixed seeds (shard index), as the reference's --seed")
    p.add_argument("--output", default="result", help="result<shard>.gfbt prefix")
    p.add_argument("--devices", type=int, default=0, help="0 = all")
    return p


def draw(args, name, rng, n):
    mean = getattr(args, "init_%s_mean" % name)
    if getattr(args, "init_%s_dist" % name) == "normal":
        return rng.normal(mean, getattr(args, "init_%s_sigma" % name), n)
    return np.full(n, mean)


def initial_conditions(args, shard, n):
    """xrays.cpp:448-453 draw order; --use_cyl_xy reads x as radius and y as angle (xrays.cpp:82-130)."""
    rng = np.random.default_rng(shard if args.seed else None)
    s = {"t": np.zeros(n)}
    for name in ("w", "kx", "ky", "kz", "z"):
        s[name] = draw(args, name, rng, n)
    if args.use_cyl_xy:
        radius = draw(args, "x", rng, n)
        phi = draw(args, "y", rng, n)
        s["x"], s["y"] = radius*np.cos(phi), radius*np.sin(phi)
    else:
        s["x"], s["y"] = draw(args, "x", rng, n), draw(args, "y", rng, n)
    return s


def trace_shard(args, shard, n, device, report):
    t0 = time.perf_counter()
    dt = args.endtime/args.num_times
    tr = RayTracer(args.dispersion, args.equilibrium, n, dt, solver=args.solver, table_file=args.equilibrium_file,
                   device=device, options="fused_steps=%d" % args.sub_steps)
    tr.set_state(initial_conditions(args, shard, n))
    solve_for = [v for v in ("kx", "ky", "kz") if getattr(args, "init_" + v)]
    tr.init(solve_for[0] if solve_for else "")
    tr.compile()
    t1 = time.perf_counter()
    records = tr.trace(max(args.num_times//args.sub_steps, 1), args.sub_steps)
    t2 = time.perf_counter()
    write_trajectory("%s%d.gfbt" % (args.output, shard), records)
    tr.close()
    report[shard] = {"rays": n, "setup_s": t1 - t0, "trace_s": t2 - t1, "records": records.shape[0],
                     "max_residual": float(np.max(records[-1][8])) if n else 0.0}


def main(argv=None):
    args = parser().parse_args(argv)
    available = lib.gfb_device_count()
    if available <= 0:
        raise SystemExit("xrays: no CUDA device (the B200 back end has no CPU fallback)")
    devices = max(1, min(args.devices or available, available, args.num_rays))
    offsets = shard_offsets(args.num_rays, devices)
    report = {}
    threads = [threading.Thread(target=trace_shard, args=(args, d, offsets[d + 1] - offsets[d], d, report))
               for d in range(devices)]
    t0 = time.perf_counter()
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    total = time.perf_counter() - t0
    steps = max(args.num_times//args.sub_steps, 1)*args.sub_steps
    slowest = max(r["trace_s"] for r in report.values())
    print("xrays: %d rays x %d steps on %d device(s): trace %.3f s (%.3e ray-steps/s incl. output copies), total %.3f s"
          % (args.num_rays, steps, devices, slowest, args.num_rays*steps/slowest, total))
    if args.verbose:
        for d in sorted(report):
            print("  shard %d: %s" % (d, report[d]))
    return 0


if __name__ == "__main__":
    raise SystemExit(main())
