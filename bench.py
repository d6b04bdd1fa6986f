#!/usr/bin/env python3
"""Headline benchmark: FP64 RK4 ray-steps/s of the EFIT ray step on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference] [--workload ...]

One bench "step" = one block of SUB_STEPS (100, the reference's sub_steps between outputs,
graph_driver/efit_example.sh) RK4 steps over every ray: one fused launch of the hot-path kernel.
`value` is ray-steps/s with the ensemble resident in HBM, timed with CUDA events on the launching
stream (max over ranks).  `e2e` is the same quantity through the public API with host buffers:
every step uploads the 8 state arrays from pinned host memory, runs the block and reads state +
residual back.  Rays are sharded over ranks with no data-path collective (weak scaling: the
per-GPU ensemble is fixed).

`--impl reference` times the reference's own CPU implementation (oracle/_ref/ref_driver: the
unmodified reference graph/solver code, kernels compiled by g++ -O3 -ffast-math) on all host
threads on a bounded sample of the same workload.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

SUB_STEPS = 100
#  From the committed ncu --set full captures of exactly this command (profiles/r1_ncu_*.txt):
#  DRAM bytes read + written per launch, and the share of cycles the FP64 pipe was busy.
NCU = {
    "efit_xmode": {"traffic": 64113152 + 16543488, "fp64_pipe_active_pct": 75.6, "source": "profiles/r1_ncu_efit_xmode_solver_kernel.txt"},
    "efit_cold": {"traffic": 64047360 + 9921792, "fp64_pipe_active_pct": 73.9, "source": "profiles/r1_ncu_efit_cold_solver_kernel.txt"},
}
WORKLOADS = {
    # name: (dispersion, equilibrium, default rays per GPU, dt)
    "efit_xmode": ("extra_ordinary_wave", "efit", 1000000, 2.0e-5),     # BASELINE configs[1]
    "efit_cold": ("cold_plasma", "efit", 1000000, 2.0e-5),              # north-star roofline kernel
    "efit_omode": ("ordinary_wave", "efit", 1000000, 2.0e-5),
    "slab_omode": ("ordinary_wave", "slab_density", 1000000, 1.0e-3),   # analytic variant of configs[0]
    "vmec_omode": ("ordinary_wave", "vmec", 1250000, 1.0e-4),           # configs[3]: 10^7 rays / 8 GPUs
    "vmec_cold": ("cold_plasma", "vmec", 1250000, 1.0e-4),
}


def generator(eq):
    from graph_framework_b200 import workloads
    return {"efit": workloads.efit_ensemble, "vmec": workloads.vmec_states}.get(eq, workloads.slab_ensemble)


def nproc():
    try:
        return len(os.sched_getaffinity(0))
    except AttributeError:
        return os.cpu_count() or 1


class ClockSampler:
    """nvidia-smi clocks and throttle reasons during the timed region (B200_PROFILING.md)."""

    def __init__(self, index):
        self.index = index
        self.rows = []
        self.proc = None

    def start(self):
        q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
             "clocks_event_reasons.sw_power_cap")
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + q,
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.time(), [c.strip() for c in line.split(",")]))

    def mark(self):
        self.t0 = time.time()

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        t0 = getattr(self, "t0", 0.0)
        for stamp, r in self.rows:
            if len(r) < 7 or stamp < t0:
                continue
            try:
                sm.append(float(r[0]))
                mx.append(float(r[1]))
            except ValueError:
                continue
            for name, v in zip(names, r[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        sm.sort()
        return {"sm_mhz": sm[len(sm)//2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "samples": len(sm), "reasons": sorted(reasons)}


def reference_arm(args, rank, world):
    """The reference's own CPU path on this box's host cores, bounded sample."""
    if rank != 0:
        return 0
    from oracle import reference
    from graph_framework_b200 import workloads
    disp, eq, _, dt = WORKLOADS[args.workload]
    cores = nproc()
    if not reference.available():
        print(json.dumps({"impl": "reference", "unavailable": "oracle/_ref/ref_driver not built"}))
        return 0
    rays = args.ref_rays
    state = generator(eq)(rays, seed=0)
    # one process: graph build, Newton init and JIT once, then W + K timed blocks of SUB_STEPS steps
    last = reference.bench(disp, eq, rays, dt, SUB_STEPS, cores, state, blocks=args.warmup + args.steps)
    times = last["block_s"][args.warmup:]
    total = sum(times)
    value = rays*SUB_STEPS*len(times)/total
    line = {
        "impl": "reference", "metric": "ray-steps/sec (FP64 RK4)", "value": value, "unit": "ray-steps/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1.0e3*total/len(times),
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": "%s: %s + %s, RK4 FP64, %d steps per block" % (args.workload, disp, eq, SUB_STEPS),
                   "rays": rays, "dt": dt},
        "cpu_baseline": {"value": value, "unit": "ray-steps/s", "cores": cores, "kind": "reference",
                         "sample": "%d rays x %d RK4 steps per timed step, reference graph+solver, kernels by g++ -O3 -ffast-math, %d threads (setup %.1fs, Newton init %.1fs, JIT %.1fs excluded as in xrays_bench.cpp)"
                                   % (rays, SUB_STEPS, cores, last["setup_s"], last["init_s"], last["compile_s"])},
        "e2e": {"value": value, "unit": "ray-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))
    return 0


def boris_arm(args, rank, local_rank, world):
    """BASELINE configs[4]: the xkorc Boris push (graph_korc/xkorc.cpp:66-121) in the EFIT field.
    Strong scaling: --rays is the TOTAL particle count, sharded over ranks."""
    import numpy as np
    import torch
    from graph_framework_b200 import workloads, parallel
    from graph_framework_b200.rays import BorisPusher
    torch.cuda.set_device(local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    total = args.rays or 100000000
    off, n = parallel.my_shard(total, rank, world)
    x, y, z, ux, uy, uz = workloads.boris_ensemble(n, seed=rank)
    push = BorisPusher("efit", n, dt=0.5, device=local_rank, options="fused_steps=%d %s" % (SUB_STEPS, args.options))
    push.set_state(x, y, z, ux, uy, uz)
    push.compile()
    # (the pusher keeps particles sorted by the (R, Z) cell of the EFIT tables, re-sorted every 1000 pushes,
    #  inside the timed region when due; --options bin_rays=0 switches it off)
    for _ in range(args.warmup):
        push.step(SUB_STEPS)
    torch.cuda.synchronize()
    if dist:
        dist.barrier()
    launches0 = push.launch_count()
    ms = 0.0
    for _ in range(args.steps):
        push.timer_start()
        push.step(SUB_STEPS)
        ms += push.timer_stop()
    launches = push.launch_count() - launches0
    torch.cuda.synchronize()
    if dist:
        dist.barrier()
        t = torch.tensor([ms], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    st = push.get_state()
    finite = bool(np.isfinite(st["gamma"]).all())
    if rank == 0:
        value = total*SUB_STEPS*args.steps/(ms*1.0e-3)
        flop = workloads.FLOP_PER_PARTICLE_STEP_BORIS
        print(json.dumps({
            "metric": "particle-steps/sec (FP64 Boris)", "value": value, "unit": "particle-steps/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms/args.steps, "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": "boris: xkorc Boris push in the EFIT field, %d fused pushes per bench step" % SUB_STEPS,
                       "particles_total": total, "dt": 0.5, "state_larger_than_L2": n*56 > 126e6, "options": args.options,
                       "binning": "particles kept sorted by EFIT (R, Z) cell, re-sorted every 1000 pushes"},
            "roofline": {"bound": "fp64", "achieved": flop*value/world/1.0e12, "peak": None, "unit": "TFLOP/s", "frac": None,
                         "traffic": None, "algorithmic_flop_per_particle_step": flop,
                         "hbm": {"algorithmic_bytes_per_launch": 112*n, "achieved_gbs": 112*n/(ms/args.steps*1.0e-3)/1.0e9}},
            "gpu_launches": launches, "finite": finite, "info": push.info()}))
    push.close()
    if dist:
        dist.barrier()
        dist.destroy_process_group()
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="efit_xmode", choices=sorted(WORKLOADS) + ["boris"])
    ap.add_argument("--rays", type=int, default=0, help="rays per GPU (default: the workload's)")
    ap.add_argument("--ref-rays", type=int, default=20000, help="rays of the bounded CPU sample")
    ap.add_argument("--options", default="", help="emit/launch options passed to gfb_rays_create")
    ap.add_argument("--chunks", type=int, default=5, help="pieces of the ensemble pipelined by the e2e call")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))

    if args.workload == "boris":
        if args.impl == "reference":
            if rank == 0:
                print(json.dumps({"impl": "reference", "unavailable": "the Boris workload has no CPU arm in bench.py (see oracle/ref_driver korc mode)"}))
            return 0
        return boris_arm(args, rank, local_rank, world)
    if args.impl == "reference":
        return reference_arm(args, rank, world)

    import numpy as np
    import torch
    from graph_framework_b200 import workloads
    from graph_framework_b200.rays import RayTracer, STATE
    from graph_framework_b200 import _lib

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the B200 back end has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    disp, eq, default_rays, dt = WORKLOADS[args.workload]
    rays = args.rays or default_rays                      # per GPU: weak scaling
    gen = generator(eq)
    state0 = gen(rays, seed=rank)

    t_setup = time.perf_counter()
    tracer = RayTracer(disp, eq, rays, dt, solver="rk4", device=local_rank,
                       options=("fused_steps=%d " % SUB_STEPS) + args.options)
    tracer.set_state(state0)
    t_init = time.perf_counter()
    tracer.init("kx")                                      # device-resident per-ray Newton
    t_compile = time.perf_counter()
    tracer.compile()
    # (tabulated equilibria: the tracer keeps rays sorted by table cell while stepping -- EFIT (R, Z) cells,
    #  VMEC radial cells; re-sorted after about half a cell of travel, inside the timed region when due;
    #  --options bin_rays=0 switches it off)
    t_ready = time.perf_counter()
    stats = tracer.kernel_stats()
    fp64_peak = tracer.fp64_peak()

    # ---- kernel-resident measurement -------------------------------------------------
    sampler = ClockSampler(local_rank)
    sampler.start()
    for _ in range(args.warmup):
        tracer.step(SUB_STEPS)
    tracer.wait()
    if dist:
        dist.barrier()
    torch.cuda.synchronize()
    sampler.mark()                                         # clocks are kept from here to the end of the e2e region
    launches0 = tracer.launch_count()
    kernel_ms = 0.0
    for _ in range(args.steps):
        tracer.flush_l2()                                  # untimed: evict state + tables from L2
        tracer.timer_start()
        tracer.step(SUB_STEPS)
        kernel_ms += tracer.timer_stop()
    tracer.wait()
    launches = tracer.launch_count() - launches0                    # solver_kernel launches (L2 fills not counted)
    torch.cuda.synchronize()
    if dist:
        dist.barrier()
        t = torch.tensor([kernel_ms], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        kernel_ms_max = float(t.item())
    else:
        kernel_ms_max = kernel_ms

    # ---- end to end through the public API with host buffers -------------------------
    host = {k: torch.empty(rays, dtype=torch.float64).pin_memory() for k in STATE + ("residual",)}
    host_out = {k: v.numpy() for k, v in host.items()}
    tracer.get_state(out=host_out)
    host_in = {k: torch.empty(rays, dtype=torch.float64).pin_memory() for k in STATE}
    for k in STATE:
        host_in[k].copy_(host[k])
    host_np = {k: host_in[k].numpy() for k in STATE}
    e2e_steps = max(2, min(args.steps, 10))
    for _ in range(2):
        out = tracer.step_host(SUB_STEPS, host_np, host_out, chunks=args.chunks)
    if dist:
        dist.barrier()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        # one public-API call: H2D of the 8 state arrays from pinned memory, SUB_STEPS fused steps,
        # D2H of the 8 arrays + residual into pinned memory; pieces of the ensemble are pipelined
        out = tracer.step_host(SUB_STEPS, host_np, host_out, chunks=args.chunks)
    torch.cuda.synchronize()
    e2e_s = time.perf_counter() - t0
    if dist:
        t = torch.tensor([e2e_s], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_s = float(t.item())
    finite = bool(np.isfinite(out["x"]).all())
    clocks = sampler.stop()

    total_rays = rays*world
    value = total_rays*SUB_STEPS*args.steps/(kernel_ms_max*1.0e-3)
    e2e_value = total_rays*SUB_STEPS*e2e_steps/e2e_s

    if rank == 0:
        flop = workloads.FLOP_PER_RAY_STEP.get((disp, eq))
        per_launch_ms = kernel_ms/args.steps
        achieved = flop*rays*SUB_STEPS/(per_launch_ms*1.0e-3)/1.0e12 if flop else None
        hbm_bytes = workloads.STATE_BYTES_PER_RAY*rays
        peaks = {}
        try:
            with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
                peaks = json.load(f)
        except OSError:
            pass
        line = {
            "metric": "ray-steps/sec (FP64 RK4)", "value": value, "unit": "ray-steps/s",
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": kernel_ms_max/args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": "%s: %s + %s, RK4 FP64, %d RK4 steps per bench step (one fused launch)"
                                   % (args.workload, disp, eq, SUB_STEPS),
                       "rays_per_gpu": rays, "dt": dt, "newton_init": "device-resident per-ray",
                       "l2": "flushed between timed steps (256 MiB fill, untimed)",
                       "options": args.options},
            "roofline": {"bound": "fp64", "achieved": achieved, "peak": fp64_peak, "unit": "TFLOP/s",
                         "frac": (achieved/fp64_peak) if achieved else None,
                         "traffic": NCU.get(args.workload, {}).get("traffic"),
                         "fp64_pipe_active_pct_ncu": NCU.get(args.workload, {}).get("fp64_pipe_active_pct"),
                         "ncu_source": NCU.get(args.workload, {}).get("source"),
                         "note": "achieved = the REFERENCE kernel's flop count per ray-step (BASELINE.md section 2) x ray-steps / time, the unit of work SURVEY.md 8d fixes; this back end executes fewer FP64 instructions for the same step (reverse-mode gradient, shared reciprocals), so frac can exceed the FP64-pipe busy share and, for cold plasma, 1.0",
                         "peak_source": "DFMA peak measured live by gfb_measure_fp64_peak (MEASURED_PEAKS.json has no FP64 figure); nominal 148 SM x 64 lanes x 2 x 1.965 GHz = 37.2",
                         "algorithmic_flop_per_ray_step": flop,
                         "hbm": {"algorithmic_bytes_per_launch": hbm_bytes,
                                 "achieved_gbs": hbm_bytes/(per_launch_ms*1.0e-3)/1.0e9,
                                 "peak_gbs": peaks.get("hbm_gbs")}},
            "e2e": {"value": e2e_value, "unit": "ray-steps/s", "h2d_bytes_per_step": 8*8*rays,
                    "d2h_bytes_per_step": 9*8*rays, "steps": e2e_steps, "finite": finite,
                    "api": "RayTracer.step_host (gfb_rays_step_host): upload, %d fused steps, read-back; %d pipelined chunks" % (SUB_STEPS, args.chunks)},
            "gpu_launches": launches,
            "kernel": dict(stats, block=128, min_blocks_per_sm=int(_lib.lib.gfb_compiled_min_blocks(tracer.ctx))),
            "clocks": clocks,
            "phases_s": {"setup": t_init - t_setup, "newton_init": t_compile - t_init, "jit": t_ready - t_compile},
        }
        if not args.no_cpu_baseline and world == 1:          # reported baseline, rank 0 at N = 1 only
            try:
                from oracle import reference
                if reference.available():
                    cores = nproc()
                    n_ref = args.ref_rays
                    ref_state = gen(n_ref, seed=0)
                    r = reference.bench(disp, eq, n_ref, dt, SUB_STEPS, cores, ref_state)
                    line["cpu_baseline"] = {
                        "value": r["ray_steps_per_s"], "unit": "ray-steps/s", "cores": cores, "kind": "reference",
                        "sample": "%d rays x %d RK4 steps, unmodified reference graph+solver, kernel compiled by g++ -O3 -ffast-math, %d threads; stepping phase only (setup %.1fs, init %.1fs, JIT %.1fs excluded)"
                                  % (n_ref, SUB_STEPS, cores, r["setup_s"], r["init_s"], r["compile_s"])}
                else:
                    line["cpu_baseline"] = {"value": None, "unit": "ray-steps/s", "cores": 0, "kind": "reference",
                                            "sample": "oracle/_ref/ref_driver not present"}
            except Exception as e:      # the baseline must never take the headline down
                line["cpu_baseline"] = {"value": None, "unit": "ray-steps/s", "cores": 0, "kind": "reference",
                                        "sample": "failed: %r" % (e,)}
        print(json.dumps(line))
    tracer.close()
    if dist:
        dist.barrier()
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
