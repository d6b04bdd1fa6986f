#!/usr/bin/env python3
"""Headline benchmark: FP64 RK4 ray-steps/s of the EFIT ray step on B200, with every BASELINE.json
configuration timed briefly beside it.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference] [--workload ...]
                    [--scaling weak|strong] [--no-extras]

One bench "step" = one output block: SUB_STEPS (100, the reference's sub_steps between outputs,
graph_driver/efit_example.sh) RK4 steps over every ray = one fused launch of the hot-path kernel.
`value` is ray-steps/s with the ensemble resident in HBM, timed with CUDA events on the launching
stream (max over ranks).  `e2e` is the same quantity through the public API with host buffers: every
step uploads the state arrays from pinned host memory, runs the block and reads the result back.
Rays are sharded over ranks with no data-path collective; the one collective of the path -- the sum
of the binned power-deposition profile, BASELINE configs[2] -- is in workload `efit_absorb`.

The headline workload (default efit_xmode = BASELINE configs[1]) fills the top-level keys; unless
--no-extras is given the other configurations are timed for 3 steps each into extra.workloads:
efit_cold (the north-star roofline kernel), efit_absorb (configs[2]), vmec_omode (configs[3]) and
boris (configs[4]).  --scaling strong shards the BASELINE totals (10^6 / 10^7 / 10^8) over the ranks.

`--impl reference` times the reference's own CPU implementation (oracle/_ref/ref_driver: the
unmodified reference graph/solver code, kernels compiled by g++ -O3 -ffast-math) on all host
threads on a bounded sample of the same workload.
"""
import argparse
import glob
import hashlib
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

SUB_STEPS = 100
WORKLOADS = {
    # name: (dispersion, equilibrium, rays per GPU (weak), dt, BASELINE total (strong))
    "efit_xmode": ("extra_ordinary_wave", "efit", 1000000, 2.0e-5, 1000000),    # BASELINE configs[1]
    "efit_cold": ("cold_plasma", "efit", 1000000, 2.0e-5, 1000000),             # north-star roofline kernel
    "efit_omode": ("ordinary_wave", "efit", 1000000, 2.0e-5, 1000000),
    "slab_omode": ("ordinary_wave", "slab_density", 1000000, 1.0e-3, 10000),    # analytic variant of configs[0]
    "vmec_omode": ("ordinary_wave", "vmec", 1250000, 1.0e-4, 10000000),         # configs[3]: 10^7 rays / 8 GPUs
    "vmec_cold": ("cold_plasma", "vmec", 1250000, 1.0e-4, 10000000),
}
#  configs[2]: O-mode rays launched at R = 2.3, just outside the electron-cyclotron resonance layer (Im k_amp
#  rises from R ~ 2.2, peaks near 2.06; tests/golden/ref_absorb_ordinary_wave_efit.npz), device Newton, weak
#  damping + power + deposition on a 64 x 64 x 128 grid over the EFIT box, one FP64 all-reduce of the block's
#  profile per block.  dt 2e-4: 0.02 of travel per block of 100 steps; after `period` blocks (R ~ 2.02, most of
#  the power absorbed; beyond the layer the weak-damping formula is not meaningful) the ensemble is put back on
#  its launch circle, untimed, so that every timed block deposits.  The beam is monochromatic (w = 700): with the
#  frequency spread of efit_example.sh the low-frequency tail of 10^6 rays passes ITS resonance layer early, and
#  behind the layer the reference's weak-damping formula returns Im k < 0, i.e. exp(+...) "absorbed power".
ABSORB = {"dispersion": "ordinary_wave", "equilibrium": "efit", "rays": 1000000, "dt": 2.0e-4, "total": 1000000,
          "radius": 2.3, "period": 13, "w": 700.0, "bins": (64, 64, 128), "lo": (0.84, -1.7, -1.6), "hi": (2.54, 1.7, 1.6)}
BORIS = {"particles": 20000000, "total": 100000000, "dt": 0.5}                  # configs[4]
EXTRAS = ("efit_cold", "efit_absorb", "vmec_omode", "boris")
EXTRA_STEPS, EXTRA_WARMUP = 3, 3


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
        return self

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


# ---------------------------------------------------------------------------------------------
#  Roofline bookkeeping
# ---------------------------------------------------------------------------------------------
def kernel_text_sha(ctx):
    """sha256 of the CUDA text NVRTC compiled for this context (skeleton + emitted bodies) and of the options
    it was given (stage unrolling, blocks/SM promise): what a committed ncu capture must have been taken from
    for its counters to describe this binary."""
    from graph_framework_b200 import _lib
    return hashlib.sha256(_lib.lib.gfb_source(ctx) + b"\n//options: " + _lib.lib.gfb_compile_options(ctx)).hexdigest()


def ncu_capture(workload, sha):
    """The committed ncu --set full capture of this workload's kernel (profiles/r*_ncu_<workload>.json,
    written on the GPU box by tools/capture_ncu.sh): newest round first.  `stale` = the kernel text
    benchmarked now is not the text the capture was taken from."""
    found = sorted(glob.glob(os.path.join(ROOT, "profiles", "r*_ncu_%s.json" % workload)), reverse=True)
    for path in found:
        try:
            with open(path) as f:
                cap = json.load(f)
        except (OSError, ValueError):
            continue
        cap["file"] = os.path.relpath(path, ROOT)
        cap["stale"] = cap.get("kernel_text_sha256") != sha
        return cap
    return None


def fp64_roofline(workload, ctx, units_per_launch, unit_steps, launch_ms, algorithmic_flop, peaks, bytes_per_launch):
    """units_per_launch: rays (particles) one launch advances; unit_steps: steps per launch.
    frac = EXECUTED FP64 flops / peak (<= 1 by construction): executed = (2 DFMA + DADD + DMUL thread
    instructions) per unit-step from the committed ncu capture of the same kernel text x the measured rate.
    work_rate_vs_reference_flops = the reference kernel's flop count per unit-step (BASELINE.md section 2,
    SURVEY.md 8d) x the measured rate / peak: how fast the REFERENCE's arithmetic is being retired; it
    exceeds frac (and may exceed 1) because this back end executes fewer flops for the same step."""
    rate = units_per_launch*unit_steps/(launch_ms*1.0e-3)           # unit-steps per second on this GPU
    peak = peaks["nominal_tflops"]
    sha = kernel_text_sha(ctx)
    cap = ncu_capture(workload, sha)
    out = {"bound": "fp64", "unit": "TFLOP/s", "peak": peak,
           "peak_source": "FP64 pipe: %d SMs x 64 lanes x 2 x %.3f GHz (device attributes); MEASURED_PEAKS.json has no FP64 figure. "
                          "The live DFMA probe (gfb_measure_fp64_peak) reads %.2f with the FP64 pipe %s busy under ncu, i.e. the same pipe peak"
                          % (peaks["sms"], peaks["clock_ghz"], peaks["probe_tflops"],
                             "92.0 % (profiles/r2_ncu_fp64_peak.json)"),
           "probe_tflops": peaks["probe_tflops"],
           "algorithmic_flop_per_unit_step": algorithmic_flop,
           "algorithmic_tflops_reference_count": algorithmic_flop*rate/1.0e12 if algorithmic_flop else None,
           "work_rate_vs_reference_flops": algorithmic_flop*rate/1.0e12/peak if algorithmic_flop else None,
           "kernel_text_sha256": sha,
           "hbm": {"algorithmic_bytes_per_launch": bytes_per_launch,
                   "achieved_gbs": bytes_per_launch/(launch_ms*1.0e-3)/1.0e9, "peak_gbs": peaks.get("hbm_gbs")}}
    if cap and cap.get("fp64_flop_per_unit_step"):
        executed = cap["fp64_flop_per_unit_step"]
        out.update({"achieved": executed*rate/1.0e12, "frac": executed*rate/1.0e12/peak,
                    "executed_flop_per_unit_step": executed,
                    "fp64_instructions_per_unit_step": cap.get("fp64_inst_per_unit_step"),
                    "fp64_pipe_active_pct_ncu": cap.get("fp64_pipe_active_pct"),
                    "traffic": cap.get("dram_bytes_per_launch"),
                    "ncu": {"file": cap["file"], "captured_at_git_sha": cap.get("git_sha"), "stale": cap["stale"],
                            "units_per_launch": cap.get("units_per_launch"), "unit_steps": cap.get("unit_steps")}})
    else:
        out.update({"achieved": out["algorithmic_tflops_reference_count"], "frac": None, "traffic": None,
                    "ncu": None, "note": "no committed ncu capture for this workload: achieved is the reference-count work rate"})
    return out


def device_peaks(tracer_like):
    """Nominal FP64 pipe peak from the device attributes + the live DFMA probe + MEASURED_PEAKS.json."""
    import torch
    props = torch.cuda.get_device_properties(torch.cuda.current_device())
    clock_khz = getattr(props, "clock_rate", None)
    if not clock_khz:
        clock_khz = 1965000
    sms = props.multi_processor_count
    peaks = {"sms": sms, "clock_ghz": clock_khz/1.0e6, "nominal_tflops": sms*64*2*clock_khz*1.0e3/1.0e12,
             "probe_tflops": tracer_like.fp64_peak()}
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            peaks["hbm_gbs"] = json.load(f).get("hbm_gbs")
    except (OSError, ValueError):
        peaks["hbm_gbs"] = None
    return peaks


class Ranks:
    """RANK / WORLD_SIZE plumbing and the max-over-ranks reduction of a device time."""

    def __init__(self):
        self.rank = int(os.environ.get("RANK", "0"))
        self.local_rank = int(os.environ.get("LOCAL_RANK", "0"))
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.dist = None

    def init(self):
        import torch
        torch.cuda.set_device(self.local_rank)
        if self.world > 1:
            import torch.distributed as dist
            dist.init_process_group("nccl", device_id=torch.device("cuda", self.local_rank))
            self.dist = dist

    def barrier(self):
        import torch
        if self.dist:
            self.dist.barrier()
        torch.cuda.synchronize()

    def max(self, value):
        if not self.dist:
            return value
        import torch
        t = torch.tensor([value], dtype=torch.float64, device="cuda")
        self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return float(t.item())

    def finish(self):
        if self.dist:
            self.dist.barrier()
            self.dist.destroy_process_group()


def shard(total, ranks):
    from graph_framework_b200 import parallel
    return parallel.my_shard(total, ranks.rank, ranks.world)[1]


# ---------------------------------------------------------------------------------------------
#  Ray workloads (EFIT / VMEC / slab step kernels)
# ---------------------------------------------------------------------------------------------
def run_rays(name, args, ranks, steps, warmup, with_e2e=True, cpu_baseline=True, rays_override=0):
    import numpy as np
    import torch
    from graph_framework_b200 import workloads, _lib
    from graph_framework_b200.rays import RayTracer, STATE

    disp, eq, per_gpu, dt, total = WORKLOADS[name]
    strong = args.scaling == "strong"
    if rays_override:
        rays = rays_override
    elif strong:
        rays = shard(total, ranks)
    else:
        rays = per_gpu
    gen = generator(eq)
    state0 = gen(rays, seed=ranks.rank)

    t_setup = time.perf_counter()
    tracer = RayTracer(disp, eq, rays, dt, solver="rk4", device=ranks.local_rank,
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
    peaks = device_peaks(tracer)

    # ---- kernel-resident measurement -------------------------------------------------
    sampler = ClockSampler(ranks.local_rank).start()
    for _ in range(warmup):
        tracer.step(SUB_STEPS)
    tracer.wait()
    ranks.barrier()
    sampler.mark()                                         # clocks are kept from here to the end of the e2e region
    launches0 = tracer.launch_count()
    kernel_ms = 0.0
    for _ in range(steps):
        tracer.flush_l2()                                  # untimed: evict state + tables from L2
        tracer.timer_start()
        tracer.step(SUB_STEPS)
        kernel_ms += tracer.timer_stop()
    tracer.wait()
    launches = tracer.launch_count() - launches0           # solver_kernel launches (+ re-sorts when due); L2 fills not counted
    ranks.barrier()
    kernel_ms_max = ranks.max(kernel_ms)

    # ---- end to end through the public API with host buffers -------------------------
    e2e = None
    finite = True
    if with_e2e:
        host = {k: torch.empty(rays, dtype=torch.float64).pin_memory() for k in STATE + ("residual",)}
        host_out = {k: v.numpy() for k, v in host.items()}
        tracer.get_state(out=host_out)
        host_in = {k: torch.empty(rays, dtype=torch.float64).pin_memory() for k in STATE}
        for k in STATE:
            host_in[k].copy_(host[k])
        host_np = {k: host_in[k].numpy() for k in STATE}
        e2e_steps = max(2, min(steps, 10))
        for _ in range(2):
            out = tracer.step_host(SUB_STEPS, host_np, host_out, chunks=args.chunks)
        ranks.barrier()
        t0 = time.perf_counter()
        for _ in range(e2e_steps):
            # one public-API call: H2D of the 8 state arrays from pinned memory, SUB_STEPS fused steps,
            # D2H of the 8 arrays + residual into pinned memory; pieces of the ensemble are pipelined
            out = tracer.step_host(SUB_STEPS, host_np, host_out, chunks=args.chunks)
        torch.cuda.synchronize()
        e2e_s = ranks.max(time.perf_counter() - t0)
        finite = bool(np.isfinite(out["x"]).all())
        copied = (8*8 + 9*8)*rays
        e2e = {"value": rays*ranks.world*SUB_STEPS*e2e_steps/e2e_s if not strong else total*SUB_STEPS*e2e_steps/e2e_s,
               "unit": "ray-steps/s", "h2d_bytes_per_step": 8*8*rays, "d2h_bytes_per_step": 9*8*rays,
               "steps": e2e_steps, "finite": finite,
               "host_copy_gbs_per_rank": copied*e2e_steps/e2e_s/1.0e9,
               "api": "RayTracer.step_host (gfb_rays_step_host): upload, %d fused steps, read-back; %d pipelined chunks"
                      % (SUB_STEPS, args.chunks)}
    clocks = sampler.stop()

    total_rays = total if strong else rays*ranks.world
    result = None
    if ranks.rank == 0:
        flop = workloads.FLOP_PER_RAY_STEP.get((disp, eq))
        result = {
            "metric": "ray-steps/sec (FP64 RK4)", "value": total_rays*SUB_STEPS*steps/(kernel_ms_max*1.0e-3),
            "unit": "ray-steps/s", "n_gpus": ranks.world, "steps": steps, "warmup": warmup,
            "ms_per_step": kernel_ms_max/steps, "higher_is_better": True, "scaling": "strong" if strong else "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": "%s: %s + %s, RK4 FP64, %d RK4 steps per bench step (one fused launch)"
                                   % (name, disp, eq, SUB_STEPS),
                       "rays_per_gpu": rays, "rays_total": total_rays, "dt": dt, "newton_init": "device-resident per-ray",
                       "l2": "flushed between timed steps (256 MiB fill, untimed)", "options": args.options},
            "roofline": fp64_roofline(name, tracer.ctx, rays, SUB_STEPS, kernel_ms/steps, flop, peaks,
                                      workloads.STATE_BYTES_PER_RAY*rays),
            "e2e": e2e, "gpu_launches": launches,
            "kernel": dict(stats, block=128, min_blocks_per_sm=int(_lib.lib.gfb_compiled_min_blocks(tracer.ctx))),
            "clocks": clocks,
            "phases_s": {"setup": t_init - t_setup, "newton_init": t_compile - t_init, "jit": t_ready - t_compile},
        }
        if cpu_baseline and ranks.world == 1:                # reported baseline, rank 0 at N = 1 only
            result["cpu_baseline"] = reference_rays_baseline(disp, eq, dt, gen, args.ref_rays)
    tracer.close()
    return result


def reference_rays_baseline(disp, eq, dt, gen, n_ref):
    try:
        from oracle import reference
        if eq == "vmec":
            return {"value": None, "unit": "ray-steps/s", "cores": 0, "kind": "reference",
                    "sample": "not timed here: the reference needs ~8 min of graph building + compiling per VMEC kernel "
                              "(oracle/make_golden.py vmec_trace); its stepping rate is 4.6e3 ray-steps/s per thread (SURVEY.md section 6)"}
        if not reference.available():
            return {"value": None, "unit": "ray-steps/s", "cores": 0, "kind": "reference", "sample": "oracle/_ref/ref_driver not present"}
        cores = nproc()
        r = reference.bench(disp, eq, n_ref, dt, SUB_STEPS, cores, gen(n_ref, seed=0))
        return {"value": r["ray_steps_per_s"], "unit": "ray-steps/s", "cores": cores, "kind": "reference",
                "sample": "%d rays x %d RK4 steps, unmodified reference graph+solver, kernel compiled by g++ -O3 -ffast-math, %d threads; "
                          "stepping phase only (setup %.1fs, init %.1fs, JIT %.1fs excluded)"
                          % (n_ref, SUB_STEPS, cores, r["setup_s"], r["init_s"], r["compile_s"])}
    except Exception as e:      # the baseline must never take the headline down
        return {"value": None, "unit": "ray-steps/s", "cores": 0, "kind": "reference", "sample": "failed: %r" % (e,)}


# ---------------------------------------------------------------------------------------------
#  configs[2]: trace + absorption + deposition, the profile summed over GPUs once per output block
# ---------------------------------------------------------------------------------------------
def run_absorb(args, ranks, steps, warmup, with_e2e=True, check=True):
    import numpy as np
    import torch
    from graph_framework_b200 import workloads
    from graph_framework_b200.rays import RayTracer, STATE, pinned_empty

    cfg = ABSORB
    strong = args.scaling == "strong"
    rays = shard(cfg["total"], ranks) if strong else cfg["rays"]
    total_rays = cfg["total"] if strong else rays*ranks.world
    bins, lo, hi = cfg["bins"], cfg["lo"], cfg["hi"]
    state0 = workloads.efit_ensemble(rays, seed=ranks.rank, radius=cfg["radius"])
    state0["w"][:] = cfg["w"]
    tracer = RayTracer(cfg["dispersion"], cfg["equilibrium"], rays, cfg["dt"], device=ranks.local_rank,
                       options=("fused_steps=%d absorption=1 " % SUB_STEPS) + args.options)
    tracer.set_state(state0)
    tracer.init("kx")
    tracer.compile()
    start = tracer.get_state(residual=False)
    peaks = device_peaks(tracer)
    #  All torch work (zeroing, NCCL) is issued on the tracer's own stream: one timeline, no host waits.
    stream = torch.cuda.ExternalStream(tracer.stream(), device=torch.device("cuda", ranks.local_rank))
    #  Two per-block histograms in flight: while block b + 1 is being traced, the NCCL all-reduce of block b runs on
    #  NCCL's own stream over NVLink; its result is added to the running profile just before its buffer is reused.
    increments = [torch.zeros(bins, dtype=torch.float64, device="cuda") for _ in range(2)]      # d_power of one block, this rank
    pending = [None, None]
    profile = torch.zeros(bins, dtype=torch.float64, device="cuda")         # running sum over blocks and ranks
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
    launch_circle = {k: pinned_empty(rays) for k in STATE}
    for k in STATE:
        launch_circle[k][:] = start[k]
    done = [0]

    def retire(i):
        """Wait (on the tracer's stream) for the reduction of buffer i and fold it into the running profile."""
        if pending[i] is not None:
            if pending[i] is not True:
                pending[i].wait()
            profile.add_(increments[i])
            pending[i] = None

    def block(timed=None):
        if done[0] % cfg["period"] == 0 and done[0]:
            tracer.put_state(launch_circle)                 # untimed: back to the launch circle, power = 1
            tracer.absorption_reset()
        i = done[0] % 2
        done[0] += 1
        with torch.cuda.stream(stream):
            if timed:
                timed[0].record(stream)
            retire(i)
            increments[i].zero_()
            tracer.deposit_block(SUB_STEPS, increments[i].data_ptr(), lo, hi, bins)
            if timed:
                timed[1].record(stream)
            if ranks.dist:
                pending[i] = ranks.dist.all_reduce(increments[i], op=ranks.dist.ReduceOp.SUM, async_op=True)   # NCCL, FP64, NVLink
            else:
                pending[i] = True
        return i

    def drain():
        with torch.cuda.stream(stream):
            for i in (done[0] % 2, (done[0] + 1) % 2):      # older buffer first
                retire(i)

    sampler = ClockSampler(ranks.local_rank).start()
    for _ in range(warmup):
        block()
    drain()
    tracer.wait()
    ranks.barrier()
    sampler.mark()
    launches0 = tracer.launch_count()
    compute_ms = 0.0
    stamps = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
    with torch.cuda.stream(stream):
        ev[2].record(stream)
    for k in range(steps):
        block(stamps[k])
    drain()
    with torch.cuda.stream(stream):
        ev[3].record(stream)
    ev[3].synchronize()
    total_ms = ev[2].elapsed_time(ev[3])
    for a, b in stamps:
        compute_ms += a.elapsed_time(b)                     # includes the wait for the reduction two blocks back, if it is late
    launches = tracer.launch_count() - launches0
    ranks.barrier()
    total_ms_max = ranks.max(total_ms)
    compute_ms_max = ranks.max(compute_ms)

    # ---- parity of the reduced profile: one more block of the production path, ALL rays -----------
    parity = None
    increment = None
    if check:
        while done[0] % cfg["period"] != cfg["period"] - 2:    # check a block in which the beam is being absorbed
            block()
        drain()
        last = block()
        with torch.cuda.stream(stream):
            if pending[last] is not True:
                pending[last].wait()
        pending[last] = None                                    # checked below, not folded into the profile twice
        tracer.wait()
        torch.cuda.synchronize()
        increment = increments[last]
        profile.add_(increment)
        pos = tracer.get_state(residual=False)                  # caller's ray order
        absorbed = tracer.get_absorbed()                        # Im k_amp, power, d_power of that block, same order
        from oracle import port
        mine = port.deposit(pos["x"], pos["y"], pos["z"], absorbed["d_power"], lo, hi, bins)
        summed = torch.from_numpy(mine).cuda()
        if ranks.dist:
            ranks.dist.all_reduce(summed, op=ranks.dist.ReduceOp.SUM)
        oracle_sum = summed.cpu().numpy()
        reduced = increment.cpu().numpy()
        scale = max(float(np.max(np.abs(oracle_sum))), 1.0e-300)
        dev = float(np.max(np.abs(reduced - oracle_sum))/scale)
        parity = {"rays_checked_per_rank": rays, "max_abs_dev_over_max_bin": dev, "tolerance": 1.0e-12, "ok": dev < 1.0e-12,
                  "oracle": "oracle.port.deposit (numpy restatement of utilities/bin.py:53-106) of every rank's rays and d_power of the block, summed over ranks",
                  "block_power": float(oracle_sum.sum()), "nonzero_bins": int((reduced != 0).sum()),
                  "rays_depositing": int((absorbed["d_power"] != 0).sum()),
                  "profile_total_all_blocks": float(profile.sum().item()),
                  "max_power_rank0": float(np.nanmax(absorbed["power"])), "min_kamp_im_rank0": float(np.nanmin(absorbed["kamp_im"])),
                  "median_transmitted_power": float(np.nanmedian(absorbed["power"]))}

    # ---- end to end: host state in, reduced profile out ------------------------------------------
    e2e = None
    if with_e2e:
        host_in = {k: pinned_empty(rays) for k in STATE}
        for k in STATE:
            host_in[k][:] = start[k]
        host_profile = torch.empty(bins, dtype=torch.float64).pin_memory()
        e2e_steps = max(2, min(steps, 5))

        def e2e_step():
            tracer.put_state(host_in)                              # H2D of the 8 state arrays
            tracer.absorption_reset()
            done[0] = 0
            i = block()
            with torch.cuda.stream(stream):
                if pending[i] is not True:
                    pending[i].wait()
                pending[i] = None
                host_profile.copy_(increments[i], non_blocking=True)   # D2H of the reduced profile of the block
            stream.synchronize()
        e2e_step()
        ranks.barrier()
        t0 = time.perf_counter()
        for _ in range(e2e_steps):
            e2e_step()
        torch.cuda.synchronize()
        e2e_s = ranks.max(time.perf_counter() - t0)
        del e2e_step, host_profile
        e2e = {"value": total_rays*SUB_STEPS*e2e_steps/e2e_s, "unit": "ray-steps/s", "h2d_bytes_per_step": 8*8*rays,
               "d2h_bytes_per_step": int(np.prod(bins))*8, "steps": e2e_steps,
               "api": "RayTracer.put_state + deposit_block + NCCL all-reduce + profile read-back"}
    clocks = sampler.stop()

    result = None
    if ranks.rank == 0:
        cells = int(np.prod(bins))
        result = {
            "metric": "ray-steps/sec (FP64 RK4 + absorption + deposition + all-reduce)",
            "value": total_rays*SUB_STEPS*steps/(total_ms_max*1.0e-3), "unit": "ray-steps/s", "n_gpus": ranks.world,
            "steps": steps, "warmup": warmup, "ms_per_step": total_ms_max/steps, "higher_is_better": True,
            "scaling": "strong" if strong else "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": "efit_absorb (BASELINE configs[2]): %s + %s, device Newton, %d RK4 steps, weak damping, power, "
                                   "deposition on %dx%dx%d bins, one NCCL FP64 all-reduce of the block's profile (%d bytes) per bench step"
                                   % (cfg["dispersion"], cfg["equilibrium"], SUB_STEPS, bins[0], bins[1], bins[2], cells*8),
                       "rays_per_gpu": rays, "rays_total": total_rays, "dt": cfg["dt"], "options": args.options,
                       "launch_radius": cfg["radius"], "restart_every_blocks": cfg["period"],
                       "l2": "state (72 MB) + absorption state (56 MB) + profile (4 MB) per GPU exceed the 126 MB L2 together; not flushed separately"},
            "collective": {"what": "torch.distributed all_reduce(SUM, float64, async) on NCCL: the reduction of block b overlaps the tracing of block b + 1 (two histograms in flight)" if ranks.dist else "none at 1 GPU",
                           "bytes": cells*8, "exposed_ms_per_step": (total_ms_max - compute_ms_max)/steps,
                           "share_of_step": (total_ms_max - compute_ms_max)/total_ms_max,
                           "note": "exposed = whole timed loop minus the trace+absorb+deposit intervals; includes zeroing and the add into the running profile (two %d-element kernels)" % cells},
            "parity": parity, "e2e": e2e, "gpu_launches": launches, "clocks": clocks,
            "roofline": fp64_roofline("efit_absorb", tracer.ctx, rays, SUB_STEPS, compute_ms/steps,
                                      workloads.FLOP_PER_RAY_STEP.get((cfg["dispersion"], cfg["equilibrium"])), peaks,
                                      workloads.STATE_BYTES_PER_RAY*rays),
        }
    #  torch frees record events on the streams a tensor was used on: every tensor that touched the tracer's
    #  stream must go before the tracer (and with it the stream) does.
    del block, drain, retire, increments, increment, pending, profile, ev, stamps, stream, launch_circle
    torch.cuda.synchronize()
    import gc
    gc.collect()
    tracer.close()
    return result


# ---------------------------------------------------------------------------------------------
#  configs[4]: the xkorc Boris push
# ---------------------------------------------------------------------------------------------
def run_boris(args, ranks, steps, warmup, particles_override=0, with_e2e=True, cpu_baseline=True):
    """BASELINE configs[4]: the xkorc Boris push (graph_korc/xkorc.cpp:66-121) in the EFIT field."""
    import numpy as np
    import torch
    from graph_framework_b200 import workloads
    from graph_framework_b200.rays import BorisPusher
    strong = args.scaling == "strong" or (args.workload == "boris" and not particles_override and not args.rays)
    if particles_override:
        n, total = particles_override, particles_override*ranks.world
        strong = False
    elif strong:
        total = args.rays or BORIS["total"]
        n = shard(total, ranks)
    else:
        n = args.rays or BORIS["particles"]
        total = n*ranks.world
    state = workloads.boris_ensemble(n, seed=ranks.rank)
    push = BorisPusher("efit", n, dt=BORIS["dt"], device=ranks.local_rank, options="fused_steps=%d %s" % (SUB_STEPS, args.options))
    push.set_state(*state)
    push.compile()
    peaks = device_peaks(push)
    # (the pusher keeps particles sorted by the (R, Z) cell of the EFIT tables, re-sorted every 1000 pushes,
    #  inside the timed region when due; --options bin_rays=0 switches it off)
    sampler = ClockSampler(ranks.local_rank).start()
    for _ in range(warmup):
        push.step(SUB_STEPS)
    ranks.barrier()
    sampler.mark()
    launches0 = push.launch_count()
    ms = 0.0
    for _ in range(steps):
        push.timer_start()
        push.step(SUB_STEPS)
        ms += push.timer_stop()
    launches = push.launch_count() - launches0
    ranks.barrier()
    ms_max = ranks.max(ms)
    e2e = None
    if with_e2e:
        e2e_steps = 2
        host = [np.array(a) for a in state]
        push.set_state(*host)
        push.step(SUB_STEPS)
        push.get_state()
        ranks.barrier()
        t0 = time.perf_counter()
        for _ in range(e2e_steps):
            push.set_state(*host)               # H2D of x, y, z, ux, uy, uz (+ the initialize_gamma pre-item)
            push.step(SUB_STEPS)
            got = push.get_state()              # D2H of the 7 arrays
        torch.cuda.synchronize()
        e2e_s = ranks.max(time.perf_counter() - t0)
        e2e = {"value": total*SUB_STEPS*e2e_steps/e2e_s, "unit": "particle-steps/s", "h2d_bytes_per_step": 6*8*n,
               "d2h_bytes_per_step": 7*8*n, "steps": e2e_steps, "finite": bool(np.isfinite(got["gamma"]).all()),
               "api": "BorisPusher.set_state + step(%d) + get_state (pageable host arrays, not pipelined)" % SUB_STEPS}
    st = push.get_state()
    finite = bool(np.isfinite(st["gamma"]).all())
    clocks = sampler.stop()
    result = None
    if ranks.rank == 0:
        flop = workloads.FLOP_PER_PARTICLE_STEP_BORIS
        result = {
            "metric": "particle-steps/sec (FP64 Boris)", "value": total*SUB_STEPS*steps/(ms_max*1.0e-3),
            "unit": "particle-steps/s", "n_gpus": ranks.world, "steps": steps, "warmup": warmup,
            "ms_per_step": ms_max/steps, "higher_is_better": True, "scaling": "strong" if strong else "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": "boris (BASELINE configs[4]): xkorc Boris push in the EFIT field, %d fused pushes per bench step" % SUB_STEPS,
                       "particles_per_gpu": n, "particles_total": total, "dt": BORIS["dt"],
                       "l2": "state larger than L2 (%d MB per GPU)" % (n*56//1000000) if n*56 > 126e6 else "state fits L2",
                       "options": args.options,
                       "binning": "particles kept sorted by EFIT (R, Z) cell, re-sorted every 1000 pushes"},
            "roofline": fp64_roofline("boris", push.ctx, n, SUB_STEPS, ms/steps, flop, peaks, 112*n),
            "e2e": e2e, "gpu_launches": launches, "finite": finite, "info": push.info(), "clocks": clocks,
        }
        if cpu_baseline and ranks.world == 1:
            result["cpu_baseline"] = reference_boris_baseline(args.ref_particles)
    push.close()
    return result


def reference_boris_baseline(n_ref):
    """The reference's own Boris work items (oracle/_ref korc mode = graph_korc/xkorc.cpp:40-121 on its CPU
    path), one process per host core as the reference runs one thread per device."""
    try:
        from oracle import reference
        from graph_framework_b200 import workloads
        if not reference.available():
            return {"value": None, "unit": "particle-steps/s", "cores": 0, "kind": "reference", "sample": "oracle/_ref/ref_driver not present"}
        cores = nproc()
        per = max(n_ref//cores, 1)
        results = [None]*cores

        def one(i):
            x, y, z, ux, uy, uz = workloads.boris_ensemble(per, seed=100 + i)
            results[i] = reference.korc("efit", x, y, z, ux, uy, uz, SUB_STEPS)[1]
        threads = [threading.Thread(target=one, args=(i,)) for i in range(cores)]
        for t in threads:
            t.start()
        for t in threads:
            t.join()
        slowest = max(r["steps_s"] for r in results)
        return {"value": per*cores*SUB_STEPS/slowest, "unit": "particle-steps/s", "cores": cores, "kind": "reference",
                "sample": "%d particles x %d pushes per process, %d processes of the unmodified reference work items (g++ -O3 -ffast-math kernel); "
                          "stepping phase only" % (per, SUB_STEPS, cores)}
    except Exception as e:
        return {"value": None, "unit": "particle-steps/s", "cores": 0, "kind": "reference", "sample": "failed: %r" % (e,)}


# ---------------------------------------------------------------------------------------------
#  The reference arm
# ---------------------------------------------------------------------------------------------
def reference_arm(args, rank, world):
    """The reference's own CPU path on this box's host cores, bounded sample."""
    if rank != 0:
        return 0
    from oracle import reference
    if not reference.available():
        print(json.dumps({"impl": "reference", "unavailable": "oracle/_ref/ref_driver not built"}))
        return 0
    cores = nproc()
    if args.workload == "boris":
        t0 = time.perf_counter()
        b = reference_boris_baseline(args.ref_particles)
        if not b["value"]:
            print(json.dumps({"impl": "reference", "unavailable": b["sample"]}))
            return 0
        line = {"impl": "reference", "metric": "particle-steps/sec (FP64 Boris)", "value": b["value"], "unit": "particle-steps/s",
                "n_gpus": args.gpus, "steps": 1, "warmup": 0, "ms_per_step": 1.0e3*(time.perf_counter() - t0),
                "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
                "config": {"workload": "boris (BASELINE configs[4]): xkorc Boris push in the EFIT field", "particles": args.ref_particles, "dt": BORIS["dt"]},
                "cpu_baseline": b, "e2e": {"value": b["value"], "unit": "particle-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
                "gpu_launches": 0}
        print(json.dumps(line))
        return 0
    name = "efit_omode" if args.workload == "efit_absorb" else args.workload
    disp, eq, _, dt, _ = WORKLOADS[name]
    if args.workload == "efit_absorb":
        dt = ABSORB["dt"]
    if eq == "vmec":
        print(json.dumps({"impl": "reference", "unavailable": "the reference needs ~8 min of graph building + compiling per VMEC kernel; not run inside bench.py"}))
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
        "config": {"workload": "%s: %s + %s, RK4 FP64, %d steps per block%s" % (args.workload, disp, eq, SUB_STEPS,
                                                                                " (stepping only: the reference's absorption stages are file-coupled)" if args.workload == "efit_absorb" else ""),
                   "rays": rays, "dt": dt},
        "cpu_baseline": {"value": value, "unit": "ray-steps/s", "cores": cores, "kind": "reference",
                         "sample": "%d rays x %d RK4 steps per timed step, reference graph+solver, kernels by g++ -O3 -ffast-math, %d threads (setup %.1fs, Newton init %.1fs, JIT %.1fs excluded as in xrays_bench.cpp)"
                                   % (rays, SUB_STEPS, cores, last["setup_s"], last["init_s"], last["compile_s"])},
        "e2e": {"value": value, "unit": "ray-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))
    return 0


def run_workload(name, args, ranks, steps, warmup, extra):
    """extra = True: a short side measurement (no CPU baseline for the ray workloads that the headline covers)."""
    if name == "boris":
        return run_boris(args, ranks, steps, warmup, particles_override=BORIS["particles"] if extra and args.scaling != "strong" else 0,
                         with_e2e=True, cpu_baseline=True)
    if name == "efit_absorb":
        return run_absorb(args, ranks, steps, warmup, with_e2e=True, check=True)
    return run_rays(name, args, ranks, steps, warmup, with_e2e=True, cpu_baseline=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="efit_xmode", choices=sorted(WORKLOADS) + ["boris", "efit_absorb"])
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"],
                    help="strong: the BASELINE totals (10^6 EFIT rays, 10^7 VMEC rays, 10^8 particles) sharded over the ranks")
    ap.add_argument("--rays", type=int, default=0, help="rays per GPU (default: the workload's); Boris: total particles")
    ap.add_argument("--ref-rays", type=int, default=20000, help="rays of the bounded CPU sample")
    ap.add_argument("--ref-particles", type=int, default=400000, help="particles of the bounded CPU sample (Boris)")
    ap.add_argument("--options", default="", help="emit/launch options passed to gfb_rays_create")
    ap.add_argument("--chunks", type=int, default=8, help="pieces of the ensemble pipelined by the e2e call")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extras", action="store_true", help="time the headline workload only")
    ap.add_argument("--no-e2e", action="store_true")
    args = ap.parse_args()

    ranks = Ranks()
    if args.impl == "reference":
        return reference_arm(args, ranks.rank, ranks.world)

    import torch
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the B200 back end has no CPU fallback")
    ranks.init()

    if args.workload == "boris":
        line = run_boris(args, ranks, args.steps, args.warmup, with_e2e=not args.no_e2e, cpu_baseline=not args.no_cpu_baseline)
    elif args.workload == "efit_absorb":
        line = run_absorb(args, ranks, args.steps, args.warmup, with_e2e=not args.no_e2e, check=True)
    else:
        override = args.rays if args.scaling == "weak" else 0
        line = run_rays(args.workload, args, ranks, args.steps, args.warmup, with_e2e=not args.no_e2e,
                        cpu_baseline=not args.no_cpu_baseline, rays_override=override)
    if not args.no_extras:
        extras = {}
        saved_options, saved_rays = args.options, args.rays
        args.rays = 0
        for name in EXTRAS:
            if name == args.workload:
                continue
            try:
                extras[name] = run_workload(name, args, ranks, EXTRA_STEPS, EXTRA_WARMUP, extra=True)
            except Exception as e:                              # a side measurement must never take the headline down
                extras[name] = {"failed": repr(e)}
                if ranks.world > 1:
                    raise
        args.options, args.rays = saved_options, saved_rays
        if ranks.rank == 0:
            line["extra"] = {"workloads": extras,
                             "note": "the other BASELINE configurations, %d timed steps after %d warm-up steps each, same contract per entry"
                                     % (EXTRA_STEPS, EXTRA_WARMUP)}
    if ranks.rank == 0:
        print(json.dumps(line))
    ranks.finish()
    return 0


if __name__ == "__main__":
    sys.exit(main())
