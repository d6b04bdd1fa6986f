"""xrays: the ray tracing stage of the reference's driver on the B200 back end.

    python -m graph_framework_b200.xrays --num_rays=100000 --num_times=100000 --sub_steps=100 \\
        --endtime=2.0 --dispersion=ordinary_wave --equilibrium=efit --equilibrium_file=efit.gfbt \\
        --init_kx --init_kx_mean=-700 --init_w_dist=normal --init_w_mean=700 --init_w_sigma=10 ... --use_cyl_xy

Option names and meaning follow /root/reference/graph_driver/xrays.cpp:955-1037 (trace_ray stage,
xrays.cpp:413-529): one host thread per device, rays split batch/extra (xrays.cpp:423-432), initial
conditions drawn per shard in the order w, kx, ky, kz, z, (x, y) (xrays.cpp:448-453), `--init_kx`
etc. select the component solved from the dispersion relation, every `sub_steps` steps a record is
written.  With --absorption_model=weak_damping the second and third stage of the reference driver
(calculate_power xrays.cpp:573-641, bin_power :674-793) and the binning of utilities/bin.py run on
the device between blocks of steps instead of re-reading the result files; result files then also
hold kamp (imaginary part), power and d_power, and bins.gfbt holds bins = sum(d_power)/num_rays with
the bin edges xbins/ybins/zbins (bin.py:20-33, 106; options --num_x --min_x --max_x ... as bin.py).
Differences: the random stream is numpy's (seeded with the shard index; the reference uses
std::mt19937_64), equilibrium files are GFBT (tools/gfbt.py converts netCDF), results are
result<shard>.gfbt with one (time, num_rays) variable per quantity, num_times/sub_steps + 1 records of
which record 0 is the initial state, as in the reference's files; the root_find absorption model
(complex root search) is not part of this back end.
"""
import argparse
import threading
import time

import os

import numpy as np

from ._lib import lib
from .rays import RayTracer, STATE, shard_offsets
from .tools.gfbt import write_gfbt, write_trajectory

VARS = ("w", "kx", "ky", "kz", "x", "y", "z")


def parser():
    p = argparse.ArgumentParser(prog="xrays", description=__doc__, formatter_class=argparse.RawDescriptionHelpFormatter)
    p.add_argument("--verbose", action="store_true")
    p.add_argument("--num_times", type=int, default=1000)
    p.add_argument("--sub_steps", type=int, default=100)
    p.add_argument("--num_rays", type=int, default=1)
    p.add_argument("--endtime", type=float, default=1.0)
    p.add_argument("--solver", default="rk4", choices=["rk2", "rk4", "split_simplextic", "adaptive_rk4"])     # xrays.cpp:312-360
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
    p.add_argument("--absorption_model", default="", choices=["", "weak_damping"])
    for a, lo, hi in (("x", 0.0, 3.0), ("y", -3.0, 3.0), ("z", -3.0, 3.0)):
        p.add_argument("--num_%s" % a, type=int, default=32)
        p.add_argument("--min_%s" % a, type=float, default=lo)
        p.add_argument("--max_%s" % a, type=float, default=hi)
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
    absorb = args.absorption_model == "weak_damping"
    tr = RayTracer(args.dispersion, args.equilibrium, n, dt, solver=args.solver, table_file=args.equilibrium_file,
                   device=device, options="fused_steps=%d absorption=%d" % (args.sub_steps, absorb))
    tr.set_state(initial_conditions(args, shard, n))
    solve_for = [v for v in ("kx", "ky", "kz") if getattr(args, "init_" + v)]
    tr.init(solve_for[0] if solve_for else "")
    tr.compile()
    t1 = time.perf_counter()
    blocks = max(args.num_times//args.sub_steps, 1)
    profile = None
#  Record 0 is the state before the first step (xrays.cpp:246-258 writes num_steps + 1 records and the
#  power stage takes record 0 as X_last): residual 0, kamp 0, power 1, d_power 0.
    start = tr.get_state(residual=False)
    first = [start[k] for k in STATE] + [np.zeros(n)]
    if absorb:
        bins = (args.num_x, args.num_y, args.num_z)
        records, absorbed, profile = tr.trace_absorb(blocks, args.sub_steps, bins=bins,
                                                     lo=(args.min_x, args.min_y, args.min_z),
                                                     hi=(args.max_x, args.max_y, args.max_z))
        records = np.concatenate([records, absorbed], axis=1)
        first += [np.zeros(n), np.ones(n), np.zeros(n)]
        names = ("t", "w", "x", "y", "z", "kx", "ky", "kz", "residual", "kamp", "power", "d_power")
    else:
        records = tr.trace(blocks, args.sub_steps)
        names = ("t", "w", "x", "y", "z", "kx", "ky", "kz", "residual")
    records = np.concatenate([np.stack(first)[None], records], axis=0)
    t2 = time.perf_counter()
    write_trajectory("%s%d.gfbt" % (args.output, shard), records, names)
    if absorb:
        report.setdefault("tracers", {})[shard] = tr       # the profile stays on the device until main() has reduced it
    else:
        tr.close()
    report[shard] = {"profile": profile,"rays": n, "setup_s": t1 - t0, "trace_s": t2 - t1, "records": records.shape[0],
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
    tracers = report.pop("tracers", {})
    slowest = max(r["trace_s"] for r in report.values())
    if args.absorption_model:
#  bin.py:106: every shard's histogram summed, divided by the total number of rays.  The shards' profiles
#  are still on their devices: one all-reduce over NVLink peer memory (gfb_allreduce_sum_f64), one read-back.
        from .parallel import allreduce_profiles_in_process
        ordered = [tracers[d] for d in sorted(tracers)]
        bins = (args.num_x, args.num_y, args.num_z)
        total_profile = allreduce_profiles_in_process(ordered).reshape(bins)/float(args.num_rays)
        for r in report.values():
            r.pop("profile")
        for tr in ordered:
            tr.close()
        write_gfbt(os.path.join(os.path.dirname(args.output), "bins.gfbt"),
                   {"bins": total_profile,
                    "xbins": np.linspace(args.min_x, args.max_x, args.num_x + 1),
                    "ybins": np.linspace(args.min_y, args.max_y, args.num_y + 1),
                    "zbins": np.linspace(args.min_z, args.max_z, args.num_z + 1)})
    else:
        for r in report.values():
            r.pop("profile")
    print("xrays: %d rays x %d steps on %d device(s): trace %.3f s (%.3e ray-steps/s incl. output copies), total %.3f s"
          % (args.num_rays, steps, devices, slowest, args.num_rays*steps/slowest, total))
    if args.verbose:
        for d in sorted(report):
            print("  shard %d: %s" % (d, report[d]))
    return 0


if __name__ == "__main__":
    raise SystemExit(main())
