"""Python view of the ray tracing hot path (include/gfb_rays.h).

Mirrors the call sequence of the reference benchmark
(/root/reference/graph_benchmark/xrays_bench.cpp:53-101):
variables -> equilibrium -> solver::rk4 -> init(kx) -> compile -> step()* -> sync_host.
All computation happens in libgfb200.so on the GPU; this module only moves
numpy arrays across the C ABI.
"""
import ctypes
import os

import numpy as np

from ._lib import lib, check, c_double_p, GfbError

STATE = ("t", "w", "x", "y", "z", "kx", "ky", "kz")
_GOLDEN = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")
_DEFAULT_EFIT = os.path.join(_GOLDEN, "efit.gfbt")
_DEFAULT_VMEC = os.path.join(_GOLDEN, "vmec.gfbt")


class _Pinned:
    """One page-locked allocation, released when the last numpy view of it is collected."""
    def __init__(self, nbytes):
        self.ptr = ctypes.c_void_p()
        check(lib.gfb_host_alloc(int(nbytes), ctypes.byref(self.ptr)), "host_alloc")

    def __del__(self):
        if getattr(self, "ptr", None) is not None and self.ptr.value:
            lib.gfb_host_free(self.ptr)
            self.ptr = None


def pinned_empty(shape, dtype=np.float64):
    """numpy array in page-locked host memory (cudaMallocHost through the C ABI): device<->host
    copies into it run at full rate and overlap with kernels."""
    count = int(np.prod(shape))
    nbytes = max(count*np.dtype(dtype).itemsize, 8)
    owner = _Pinned(nbytes)
    buffer = (ctypes.c_byte*nbytes).from_address(owner.ptr.value)
    buffer._owner = owner               # numpy keeps `buffer` alive through .base, and with it the owner
    return np.frombuffer(buffer, dtype=dtype, count=count).reshape(shape)


def _ptr_array(arrays, n):
    arr_t = c_double_p * n
    return arr_t(*[a.ctypes.data_as(c_double_p) if a is not None else None for a in arrays])


class RayTracer:
    """solver::rk4<dispersion::X> on one GPU (solver.hpp:677-870)."""

    def __init__(self, dispersion, equilibrium, num_rays, dt, solver="rk4", table_file=None,
                 device=0, options=None):
        if equilibrium == "efit" and table_file is None:
            table_file = _DEFAULT_EFIT
        if equilibrium == "vmec" and table_file is None:
            table_file = _DEFAULT_VMEC
        self.n = int(num_rays)
        self.h = lib.gfb_rays_create(dispersion.encode(), equilibrium.encode(),
                                     (table_file or "").encode(), solver.encode(), self.n,
                                     float(dt), int(device), options.encode() if options else None)
        if not self.h:
            raise GfbError("gfb_rays_create failed: %s" % lib.gfb_last_error().decode())
        self.ctx = lib.gfb_rays_ctx(self.h)

    def close(self):
        if getattr(self, "h", None):
            lib.gfb_rays_destroy(self.h)
            self.h = None

    def __del__(self):
        self.close()

    # -- state ---------------------------------------------------------------
    def _as_arrays(self, state):
        out = []
        for name in STATE:
            v = state.get(name) if isinstance(state, dict) else state[STATE.index(name)]
            if v is None:
                out.append(None)
                continue
            a = np.ascontiguousarray(np.broadcast_to(np.asarray(v, dtype=np.float64), (self.n,)))
            out.append(a)
        return out

    def set_state(self, state):
        """variable->set(...) (xrays_bench.cpp:62-71); dict of name->array/scalar or list in STATE order."""
        arrs = self._as_arrays(state)
        check(lib.gfb_rays_set_state(self.h, _ptr_array(arrs, 8)), "set_state")

    def put_state(self, state):
        """Host->device copy of state arrays (solver_interface::sync_device)."""
        arrs = self._as_arrays(state)
        check(lib.gfb_rays_put_state(self.h, _ptr_array(arrs, 8)), "put_state")

    def init(self, var="kx", tolerance=1.0e-30, max_iterations=1000, mode="per_ray"):
        """solver_interface::init (solver.hpp:254-274): Newton solve of D = 0 for `var`."""
        check(lib.gfb_rays_init(self.h, (var or "").encode(), tolerance, max_iterations,
                                0 if mode == "per_ray" else 1), "init")

    def compile(self):
        check(lib.gfb_rays_compile(self.h), "compile")

    def step(self, num_steps=1):
        check(lib.gfb_rays_step(self.h, int(num_steps)), "step")

    def wait(self):
        check(lib.gfb_rays_wait(self.h), "wait")

    def get_state(self, residual=True, out=None):
        """sync_host (solver.hpp:368-377): returns dict of arrays (+ 'residual').
        `out` may hold preallocated (e.g. pinned) arrays keyed like the result."""
        arrs = [out[k] if out is not None and k in out else np.empty(self.n, dtype=np.float64) for k in STATE]
        res = None
        if residual:
            res = out["residual"] if out is not None and "residual" in out else np.empty(self.n, dtype=np.float64)
        check(lib.gfb_rays_get_state(self.h, _ptr_array(arrs, 8),
                                     res.ctypes.data_as(c_double_p) if residual else None), "get_state")
        out = dict(zip(STATE, arrs))
        if residual:
            out["residual"] = res
        return out

    def step_host(self, num_steps, state_in, state_out, chunks=8):
        """sync_device, num_steps steps and sync_host as one pipelined call (solver.hpp:354-384):
        upload, stepping and read-back of `chunks` pieces of the ensemble overlap.  state_in: dict
        of host arrays (pinned for full speed); state_out: dict of preallocated arrays, may
        include 'residual'."""
        ins = [state_in.get(k) for k in STATE]
        outs = [state_out.get(k) for k in STATE]
        res = state_out.get("residual")
        check(lib.gfb_rays_step_host(self.h, int(num_steps), _ptr_array(ins, 8), _ptr_array(outs, 8),
                                     res.ctypes.data_as(c_double_p) if res is not None else None, int(chunks)), "step_host")
        return state_out

    def trace(self, num_blocks, sub_steps, out=None):
        """The reference's output loop (xrays.cpp:246-259: sub_steps steps, then write_step):
        returns an array [num_blocks, 9, num_rays] with rows t, w, x, y, z, kx, ky, kz, residual.
        The copy of block b overlaps the stepping of block b + 1; `out` may be a preallocated
        (pinned) array of that shape."""
        if out is None:
            out = pinned_empty((num_blocks, 9, self.n))
        assert out.shape == (num_blocks, 9, self.n) and out.dtype == np.float64 and out.flags.c_contiguous
        check(lib.gfb_rays_trace(self.h, int(num_blocks), int(sub_steps), out.ctypes.data_as(c_double_p)), "trace")
        return out

    def trace_absorb(self, num_blocks, sub_steps, bins=None, lo=None, hi=None, profile=None, records=True,
                     records_out=None, absorbed_out=None):
        """Trace with power absorption (tracer created with options "absorption=1"): after every
        block of sub_steps steps the weak-damping and power kernels of the reference driver's
        second and third stage (absorption.hpp:395-412, xrays.cpp:693-736) run on the state in
        device memory.  Returns (records [blocks, 9, n] or None, absorbed [blocks, 3, n] with rows
        Im k_amp, power, d_power, profile [bins] or None).  `profile` continues an earlier one;
        `records_out` / `absorbed_out` are preallocated (pinned) arrays of those shapes."""
        rec = None
        if records:
            rec = records_out if records_out is not None else pinned_empty((num_blocks, 9, self.n))
            assert rec.shape == (num_blocks, 9, self.n) and rec.dtype == np.float64 and rec.flags.c_contiguous
        absorbed = absorbed_out if absorbed_out is not None else pinned_empty((num_blocks, 3, self.n))
        assert absorbed.shape == (num_blocks, 3, self.n) and absorbed.dtype == np.float64 and absorbed.flags.c_contiguous
        args = [None, None, None, None]
        if bins is not None:
            bins = tuple(int(b) for b in bins)
            profile = np.zeros(bins, dtype=np.float64) if profile is None else np.ascontiguousarray(profile, dtype=np.float64)
            assert profile.shape == bins
            lo_a, hi_a = np.asarray(lo, dtype=np.float64), np.asarray(hi, dtype=np.float64)
            bins_a = (ctypes.c_int*3)(*bins)
            args = [profile.ctypes.data_as(c_double_p), lo_a.ctypes.data_as(c_double_p),
                    hi_a.ctypes.data_as(c_double_p), bins_a]
        else:
            profile = None
        check(lib.gfb_rays_trace_absorb(self.h, int(num_blocks), int(sub_steps),
                                        rec.ctypes.data_as(c_double_p) if records else None,
                                        absorbed.ctypes.data_as(c_double_p), *args), "trace_absorb")
        return rec, absorbed, profile

    def deposit_block(self, sub_steps, profile_ptr, lo, hi, bins):
        """One output block with the deposition profile resident on the device (gfb_rays_deposit_block):
        sub_steps RK steps, weak damping, power, d_power added into the device array at `profile_ptr`
        (bins[0]*bins[1]*bins[2] doubles, e.g. tensor.data_ptr()) on the tracer's stream.  Nothing is
        copied to the host and the call does not wait."""
        lo_a = (ctypes.c_double*3)(*lo)
        hi_a = (ctypes.c_double*3)(*hi)
        bins_a = (ctypes.c_int*3)(*[int(b) for b in bins])
        check(lib.gfb_rays_deposit_block(self.h, int(sub_steps), ctypes.c_void_p(int(profile_ptr)), lo_a, hi_a, bins_a),
              "deposit_block")

    def get_dt(self):
        """solver="adaptive_rk4": the per-ray step the solver chose last."""
        out = np.empty(self.n, dtype=np.float64)
        check(lib.gfb_rays_get_dt(self.h, out.ctypes.data_as(c_double_p)), "get_dt")
        return out

    def get_absorbed(self):
        """Im k_amp, power and d_power of the last absorption block, in the caller's ray order."""
        arrs = [np.empty(self.n, dtype=np.float64) for _ in range(3)]
        check(lib.gfb_rays_get_absorbed(self.h, _ptr_array(arrs, 3)), "get_absorbed")
        return dict(zip(("kamp_im", "power", "d_power"), arrs))

    def profile_key(self):
        """(buffer key, cells) of the device-resident profile of the last trace_absorb."""
        key, cells = ctypes.c_uint64(0), ctypes.c_size_t(0)
        check(lib.gfb_rays_profile(self.h, ctypes.byref(key), ctypes.byref(cells)), "profile")
        return key.value, cells.value

    def stream(self):
        """The tracer's cudaStream_t (an integer), e.g. for torch.cuda.ExternalStream."""
        return int(lib.gfb_stream(self.ctx) or 0)

    def set_binning(self, name, lo, hi, cells, rebin_every=0):
        """Keep rays sorted by the table cell of state `name` while stepping (include/gfb_rays.h
        gfb_rays_set_binning); invisible to the caller.  name=None switches it off."""
        which = -1 if name is None else STATE.index(name)
        check(lib.gfb_rays_set_binning(self.h, which, float(lo), float(hi), int(cells), int(rebin_every)), "set_binning")

    def absorption_reset(self):
        check(lib.gfb_rays_absorption_reset(self.h), "absorption_reset")

    def rhs(self):
        """dx/dt, dy/dt, dz/dt, dkx/dt, dky/dt, dkz/dt, D at the current host state."""
        arrs = [np.empty(self.n, dtype=np.float64) for _ in range(7)]
        check(lib.gfb_rays_rhs(self.h, _ptr_array(arrs, 7)), "rhs")
        return dict(zip(("dxdt", "dydt", "dzdt", "dkxdt", "dkydt", "dkzdt", "D"), arrs))

    def device_ptr(self, name):
        which = 8 if name == "residual" else STATE.index(name)
        p = ctypes.c_void_p()
        check(lib.gfb_rays_device_ptr(self.h, which, ctypes.byref(p)), "device_ptr")
        return p.value

    def source(self):
        return lib.gfb_rays_source(self.h).decode()

    def kernel_stats(self):
        v = [ctypes.c_int(0) for _ in range(6)]
        check(lib.gfb_rays_kernel_stats(self.h, *[ctypes.byref(x) for x in v]), "kernel_stats")
        keys = ("statements", "divides", "reciprocals", "registers", "local_bytes", "smem_bytes")
        return dict(zip(keys, [x.value for x in v]))

    # -- device helpers --------------------------------------------------------
    def timer_start(self):
        check(lib.gfb_timer_start(self.ctx), "timer_start")

    def timer_stop(self):
        ms = ctypes.c_float(0)
        check(lib.gfb_timer_stop(self.ctx, ctypes.byref(ms)), "timer_stop")
        return ms.value

    def launch_count(self):
        return int(lib.gfb_launch_count(self.ctx))

    def flush_l2(self):
        check(lib.gfb_flush_l2(self.ctx), "flush_l2")

    def fp64_peak(self):
        t = ctypes.c_double(0)
        ms = ctypes.c_float(0)
        check(lib.gfb_measure_fp64_peak(self.ctx, ctypes.byref(t), ctypes.byref(ms)), "fp64_peak")
        return t.value


class BorisPusher:
    """The xkorc step graph (graph_korc/xkorc.cpp:40-121) on one GPU."""

    fp64_peak = RayTracer.fp64_peak
    flush_l2 = RayTracer.flush_l2

    NAMES = ("x", "y", "z", "ux", "uy", "uz", "gamma")

    def __init__(self, equilibrium, num_particles, dt=0.5, table_file=None, device=0, options=None):
        if equilibrium == "efit" and table_file is None:
            table_file = _DEFAULT_EFIT
        self.n = int(num_particles)
        self.h = lib.gfb_boris_create(equilibrium.encode(), (table_file or "").encode(), self.n, float(dt),
                                      int(device), options.encode() if options else None)
        if not self.h:
            raise GfbError("gfb_boris_create failed: %s" % lib.gfb_last_error().decode())
        self.ctx = lib.gfb_boris_ctx(self.h)

    def close(self):
        if getattr(self, "h", None):
            lib.gfb_boris_destroy(self.h)
            self.h = None

    def __del__(self):
        self.close()

    def set_state(self, x, y, z, ux, uy, uz):
        arrs = [np.ascontiguousarray(np.broadcast_to(np.asarray(v, dtype=np.float64), (self.n,)))
                for v in (x, y, z, ux, uy, uz)]
        check(lib.gfb_boris_set_state(self.h, _ptr_array(arrs, 6)), "boris set_state")

    def compile(self):
        check(lib.gfb_boris_compile(self.h), "boris compile")

    def set_binning(self, r_grid, z_grid, rebin_every=100):
        """Keep particles sorted by the (R, Z) cell of the field tables while stepping; invisible to
        the caller.  r_grid, z_grid = (lo, hi, cells); None switches it off."""
        if r_grid is None:
            check(lib.gfb_boris_set_binning(self.h, None, None, None, 0), "boris set_binning")
            return
        lo = np.array([r_grid[0], z_grid[0]], dtype=np.float64)
        hi = np.array([r_grid[1], z_grid[1]], dtype=np.float64)
        cells = (ctypes.c_uint*2)(int(r_grid[2]), int(z_grid[2]))
        check(lib.gfb_boris_set_binning(self.h, lo.ctypes.data_as(c_double_p), hi.ctypes.data_as(c_double_p), cells,
                                        int(rebin_every)), "boris set_binning")

    def step(self, n=1):
        check(lib.gfb_boris_step(self.h, int(n)), "boris step")

    def get_state(self):
        arrs = [np.empty(self.n, dtype=np.float64) for _ in range(7)]
        check(lib.gfb_boris_get_state(self.h, _ptr_array(arrs, 7)), "boris get_state")
        return dict(zip(self.NAMES, arrs))

    def info(self):
        b0 = ctypes.c_double(0)
        rl = ctypes.c_double(0)
        check(lib.gfb_boris_info(self.h, ctypes.byref(b0), ctypes.byref(rl)), "boris info")
        return {"b0": b0.value, "larmor_radius": rl.value}

    def timer_start(self):
        check(lib.gfb_timer_start(self.ctx), "timer_start")

    def timer_stop(self):
        ms = ctypes.c_float(0)
        check(lib.gfb_timer_stop(self.ctx, ctypes.byref(ms)), "timer_stop")
        return ms.value

    def launch_count(self):
        return int(lib.gfb_launch_count(self.ctx))


def shard_sizes(total, shards):
    """The reference's split: batch = N/G, first N%G shards get one more (xrays.cpp:423-432)."""
    batch, extra = divmod(int(total), int(shards))
    return [batch + (1 if g < extra else 0) for g in range(shards)]


def shard_offsets(total, shards):
    sizes = shard_sizes(total, shards)
    offs = [0]
    for s in sizes:
        offs.append(offs[-1] + s)
    return offs
