"""Multi-GPU plumbing: one process per GPU, rays sharded with no data-path collective.

The reference runs one std::thread per device, each with its own rays, and never exchanges
data (graph_driver/xrays.cpp:419-527).  Here the unit is one process per GPU launched by
torchrun; torch.distributed (NCCL over NVLink on the GPU box, gloo in CPU tests) is used for
exactly one thing: summing the binned power-deposition profile across ranks.
"""
import os

import numpy as np

from .rays import shard_offsets, shard_sizes     # noqa: F401  (re-exported)


def rank_world():
    return int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))


def my_shard(total, rank=None, world=None):
    """(offset, size) of this rank's contiguous slice, the reference's batch/extra split
    (xrays.cpp:423-432, xrays_bench.cpp:38-51)."""
    r, w = rank_world()
    rank = r if rank is None else rank
    world = w if world is None else world
    offs = shard_offsets(total, world)
    return offs[rank], offs[rank + 1] - offs[rank]


def shard_state(state, rank=None, world=None):
    n = len(next(iter(state.values())))
    off, size = my_shard(n, rank, world)
    return {k: np.ascontiguousarray(v[off:off + size]) for k, v in state.items()}


def allreduce_profile(hist, total_rays=None):
    """Sum a deposition histogram (torch tensor, FP64) over all ranks in place and, like
    utilities/bin.py:96-106, normalise by the total number of rays when given.
    One collective per output block; nothing else in the path communicates."""
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(hist, op=dist.ReduceOp.SUM)
    if total_rays:
        hist /= float(total_rays)
    return hist


class OverlappedProfileReducer:
    """One process per GPU: per-block deposition histograms summed over ranks WITHOUT stalling the trace.

    Two histograms are in flight.  Usage per output block:
        buf = reducer.begin_block()      # a zeroed histogram to deposit into (RayTracer.deposit_block(..., buf.data_ptr(), ...))
        reducer.end_block()              # starts the asynchronous all-reduce of that histogram (NCCL over NVLink / gloo)
    The all-reduce of block b runs while block b + 1 is traced; its result is folded into `profile` just before
    its buffer is handed out again, or by `finish()`.  All tensor work is issued on the CURRENT torch stream: on
    the GPU wrap the calls in `with torch.cuda.stream(torch.cuda.ExternalStream(tracer.stream()))` so that the
    tracer's kernels, the zeroing and the collective share one timeline (bench.py --workload efit_absorb).
    With a single rank (or torch.distributed not initialised) the blocks are simply accumulated."""

    def __init__(self, bins, device="cpu", total_rays=None):
        import torch
        self._torch = torch
        self.buffers = [torch.zeros(tuple(bins), dtype=torch.float64, device=device) for _ in range(2)]
        self.pending = [None, None]
        self.profile = torch.zeros(tuple(bins), dtype=torch.float64, device=device)
        self.total_rays = total_rays
        self.blocks = 0

    def _distributed(self):
        import torch.distributed as dist
        return dist if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1 else None

    def _retire(self, i):
        if self.pending[i] is not None:
            if self.pending[i] is not True:
                self.pending[i].wait()
            self.profile.add_(self.buffers[i])
            self.pending[i] = None

    def begin_block(self):
        i = self.blocks % 2
        self._retire(i)
        self.buffers[i].zero_()
        return self.buffers[i]

    def end_block(self):
        i = self.blocks % 2
        dist = self._distributed()
        self.pending[i] = dist.all_reduce(self.buffers[i], op=dist.ReduceOp.SUM, async_op=True) if dist else True
        self.blocks += 1

    def finish(self):
        """Folds the blocks still in flight (older first) and returns the profile, divided by total_rays when
        given (utilities/bin.py:96-106)."""
        for i in (self.blocks % 2, (self.blocks + 1) % 2):
            self._retire(i)
        return self.profile/float(self.total_rays) if self.total_rays else self.profile


def allreduce_profiles_in_process(tracers):
    """Thread-per-device model (one process, one RayTracer per GPU, the reference's xrays.cpp:419-527):
    sums the device-resident deposition profiles of all tracers in place over NVLink peer memory
    (gfb_allreduce_sum_f64) and returns the summed profile read back from the first device.
    Call when no other thread is using the tracers."""
    import ctypes
    from ._lib import lib, check
    keys_cells = [t.profile_key() for t in tracers]
    cells = keys_cells[0][1]
    assert all(c == cells for _, c in keys_cells), "profiles differ in size"
    ctxs = (ctypes.c_void_p*len(tracers))(*[t.ctx for t in tracers])
    keys = (ctypes.c_uint64*len(tracers))(*[k for k, _ in keys_cells])
    check(lib.gfb_allreduce_sum_f64(ctxs, len(tracers), keys, cells), "allreduce_sum_f64")
    out = np.empty(cells, dtype=np.float64)
    check(lib.gfb_copy_d2h(tracers[0].ctx, keys_cells[0][0], out.ctypes.data_as(ctypes.c_void_p), out.nbytes), "profile d2h")
    return out


def deposit(tracer, weight, hist, lo, hi):
    """Bin per-ray weights at the rays' current positions into `hist` (a contiguous FP64 CUDA
    tensor of shape bins) on the tracer's stream (kernels.cu deposit_kernel).
    `weight` is a CUDA tensor or device pointer of num_rays doubles."""
    import ctypes
    from ._lib import lib, check
    bins = (ctypes.c_int*3)(*hist.shape)
    lo_a = (ctypes.c_double*3)(*lo)
    hi_a = (ctypes.c_double*3)(*hi)
    wptr = weight if isinstance(weight, int) else weight.data_ptr()
    check(lib.gfb_deposit(tracer.ctx, tracer.device_ptr("x"), tracer.device_ptr("y"), tracer.device_ptr("z"),
                          wptr, tracer.n, hist.data_ptr(), lo_a, hi_a, bins), "deposit")
    check(lib.gfb_wait(tracer.ctx), "deposit wait")
    return hist
