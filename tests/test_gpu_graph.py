"""Back-end contract tests through the C binding on the GPU, modelled on the reference's
graph_tests/jit_test.cpp (kernel == host evaluate for every operator and derivative),
workflow_test.cpp (setters, repeated items), piecewise_test.cpp (table kernels) and the
converge item of c_binding_test.c."""
import ctypes

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture()
def g(lib):
    from graph_framework_b200.graph import Context
    c = Context()
    yield c
    c.close()


def test_every_operator_kernel_matches_host_evaluate(g):
    """jit_test.cpp:49-70, 165-340: add_kernel -> compile -> run -> copy_to_host vs node->evaluate()."""
    n = 257
    rng = np.random.default_rng(0)
    xs, ys, zs = rng.uniform(0.2, 3.0, n), rng.uniform(0.2, 3.0, n), rng.uniform(-2.0, 2.0, n)
    x, y, z = g.variable(n, "x", xs), g.variable(n, "y", ys), g.variable(n, "z", zs)
    exprs = {
        "add": x + y, "sub": x - z, "mul": x*z, "div": z/y, "fma": g.fma(x, y, z),
        "sqrt": g.sqrt(x), "exp": g.exp(z), "log": g.log(y), "pow": g.pow(x, y), "pow15": g.pow(x, 1.5),
        "sin": g.sin(z), "cos": g.cos(z), "atan": g.atan(x, z), "chain": g.sqrt(x*x + y*y)/(1.0 + g.exp(-z)),
    }
    outs = []
    for name, e in exprs.items():
        outs += [e, e.df(x), e.df(y), e.df(z)]
    outs = [o for o in outs if o.evaluate().size == n]          # constants are not kernel outputs
    g.add_item([x, y, z], outs, [], "operators", n)
    g.compile()
    g.run()
    for o in outs:
        got = g.copy_to_host(o, n)
        ref = o.evaluate()
        # rcp/rsqrt seeds are refined to <= 1 ulp, libm calls differ by a few ulp
        assert np.allclose(got, ref, rtol=4.0e-15, atol=1.0e-300), np.max(np.abs(got - ref)/np.abs(ref))


def test_division_and_sqrt_edge_values(g):
    """Refined-seed reciprocal / rsqrt at the edges: sqrt(0) = 0, sqrt(inf) = inf and 1/sqrt keep
    the IEEE results (k = 0 cut-off searches, physics_test.cpp:380-470, evaluate sqrt(0)); a plain
    x/0 or x/inf is allowed to be NaN (skeleton.cuh explains the trade)."""
    vals = np.array([0.0, 1.0e-300, 1.0, 4.0, 1.0e300, np.inf])
    n = vals.size
    x = g.variable(n, "x", vals)
    one = g.variable(n, "one", np.ones(n))
    outs = [g.sqrt(x), one/x, g.sqrt(x)*g.exp(x - x), one/g.sqrt(x)]
    g.add_item([x, one], outs, [], "edges", n)
    g.compile()
    g.run()
    with np.errstate(all="ignore"):
        ref = [np.sqrt(vals), 1.0/vals, np.sqrt(vals), 1.0/np.sqrt(vals)]
    for j, (o, r) in enumerate(zip(outs, ref)):
        got = g.copy_to_host(o, n)
        assert np.allclose(got[1:5], r[1:5], rtol=4.0e-16, atol=0.0), (j, got, r)
        if j != 1:
            assert got[0] == r[0] or (np.isinf(got[0]) and np.isinf(r[0])), (j, got, r)
    assert g.copy_to_host(outs[0], n)[5] == np.inf


def test_reciprocal_and_sqrt_accuracy(g):
    """gfb::rcp / gfb::rsqrt / sqrt_from_rsqrt over 20 decades: within 1 ulp of the IEEE result."""
    rng = np.random.default_rng(7)
    n = 200000
    vals = rng.uniform(1.0, 10.0, n)*10.0**rng.integers(-150, 150, n)*rng.choice([-1.0, 1.0], n)
    x = g.variable(n, "x", vals)
    a = g.variable(n, "a", np.abs(vals))
    one = g.variable(n, "one", np.ones(n))
    outs = [one/x, g.sqrt(a), one/g.sqrt(a)]
    g.add_item([x, a, one], outs, [], "accuracy", n)
    g.compile()
    g.run()
    ref = [1.0/vals, np.sqrt(np.abs(vals)), 1.0/np.sqrt(np.abs(vals))]
    ulp = 2.0**-52
    for o, r, bound in zip(outs, ref, (1.0, 1.0, 1.5)):
        got = g.copy_to_host(o, n)
        err = np.max(np.abs(got - r)/np.abs(r))
        assert err <= bound*ulp, err/ulp


def test_workflow_setters_and_repeated_items(g):
    """workflow_test.cpp:19-70: several setters read the OLD values; ten runs fuse into one launch."""
    n = 100
    a = g.variable(n, "a", np.full(n, 2.0))
    b = g.variable(n, "b", np.arange(n, dtype=float))
    c = g.variable(n, "c", np.ones(n))
    g.add_item([a, b, c], [a*b], [(a + 1.0, a), (b + a, b), (c*2.0, c)], "update", n)
    g.compile()
    for _ in range(10):
        g.run()
    av, bv, cv = g.copy_to_host(a, n), g.copy_to_host(b, n), g.copy_to_host(c, n)
    ea, eb, ec = np.full(n, 2.0), np.arange(n, dtype=float), np.ones(n)
    last = None
    for _ in range(10):
        last = ea*eb
        ea, eb, ec = ea + 1.0, eb + ea, ec*2.0
    assert np.array_equal(av, ea) and np.array_equal(bv, eb) and np.array_equal(cv, ec)
    assert last is not None


def test_output_holds_value_of_last_step(g):
    n = 64
    a = g.variable(n, "a", np.full(n, 3.0))
    out = a*a
    g.add_item([a], [out], [(a + 1.0, a)], "square", n)
    g.compile()
    for _ in range(5):
        g.run()
    assert np.array_equal(g.copy_to_host(a, n), np.full(n, 8.0))
    assert np.array_equal(g.copy_to_host(out, n), np.full(n, 49.0))      # computed from a = 7 in the fifth step


def test_converge_item_newton_sqrt(g):
    """c_binding_test.c converge item: Newton for x^2 = s with the reference's stopping rule."""
    n = 33
    s = np.linspace(1.0, 50.0, n)
    sv = g.variable(n, "s", s)
    x = g.variable(n, "x", np.ones(n))
    f = x*x - sv
    g.add_converge_item([sv, x], [f*f], [(x - f/f.df(x), x)], "newton_sqrt", n, 1.0e-30, 1000)
    g.compile()
    g.run()
    assert np.allclose(g.copy_to_host(x, n), np.sqrt(s), rtol=1.0e-15)


def test_piecewise_kernels_match_host(g):
    """piecewise_test.cpp:53-75: table kernels (global and TMA-staged groups) vs host evaluation,
    including arguments outside the grid (clamped to the edge cells)."""
    n = 500
    rng = np.random.default_rng(1)
    xs = rng.uniform(-0.3, 1.3, n)
    ys = rng.uniform(-0.3, 1.3, n)
    x, y = g.variable(n, "x", xs), g.variable(n, "y", ys)
    t1a, t1b = rng.normal(size=40), rng.normal(size=40)
    t2a, t2b = rng.normal(size=(30, 20)), rng.normal(size=(30, 20))
    p1a = g.piecewise_1D(x, 1.0/40, 0.0, t1a)
    p1b = g.piecewise_1D(x, 1.0/40, 0.0, t1b)
    p2a = g.piecewise_2D(20, x, 1.0/30, 0.0, y, 1.0/20, 0.0, t2a.ravel())
    p2b = g.piecewise_2D(20, x, 1.0/30, 0.0, y, 1.0/20, 0.0, t2b.ravel())
    outs = [p1a*x + p1b, p2a + p2b*y, (p1a + p2a)*(p1b - p2b)]
    g.add_item([x, y], outs, [], "tables", n)
    g.compile()
    g.run()
    for o in outs:
        assert np.allclose(g.copy_to_host(o, n), o.evaluate(), rtol=1.0e-13, atol=1.0e-14)   # device contracts a*b + c into one FMA
    src = g.source()
    assert "gfb::smem" in src and "__ldg" in src


@pytest.mark.parametrize("fast", [True, False])
def test_cell_selection_at_cell_edges_matches_the_reference(lib, fast):
    """Index work must be bit-exact.  7925 arguments on, and 1-3 ulp either side of, every cell edge of the
    EFIT R, Z and psi grids (plus random ones): the default kernels pick exactly the cell the reference's own
    compiled kernels pick (tests/golden/ref_cells_efit.npz from `ref_driver cells`; under -ffast-math they
    multiply by 1/scale and differ from the written rule in 78 of these arguments); with
    graph_set_fast_division(false) the kernels follow (x - offset)/scale as piecewise.hpp:26-65 is written."""
    from graph_framework_b200.graph import Context
    from conftest import golden
    g = golden("ref_cells_efit")
    differ = 0
    for tag in ("r", "z", "psi"):
        x = g["x_" + tag]
        scale, off, n = g["grid_" + tag]
        n = int(n)
        c = Context()
        c.set_fast_division(fast)
        xv = c.variable(x.size, "x", x)
        cell = c.piecewise_1D(xv, scale, off, np.arange(n, dtype=np.float64))
        c.add_item([xv], [cell], [], "cells", x.size)
        c.compile()
        c.run()
        got = c.copy_to_host(cell, x.size)
        c.close()
        written = np.trunc(np.clip((x - off)/scale, 0, n - 1))
        differ += int((g["cell_" + tag] != written).sum())
        assert np.array_equal(got, g["cell_" + tag] if fast else written), (tag, fast)
    assert differ >= 50          # the two rules are distinguishable on this set


def test_max_reduction_handles_sign_and_size(lib):
    """cuda_context.hpp:954-995 replacement: grid-wide maximum, any sign, ragged sizes."""
    ctx = lib.gfb_ctx_create(0)
    assert ctx
    rng = np.random.default_rng(2)
    for n in (1, 31, 1000, 1 << 20, (1 << 20) + 17):
        for shift in (0.0, -10.0):
            a = rng.normal(size=n) + shift
            key = 1000 + n + int(shift)
            assert lib.gfb_buffer(ctx, key, a.nbytes, a.ctypes.data_as(ctypes.c_void_p), None) == 0
            out = ctypes.c_double(0)
            assert lib.gfb_max(ctx, key, n, ctypes.byref(out)) == 0
            assert out.value == a.max()
    lib.gfb_ctx_destroy(ctx)


def test_max_reduction_ignores_nan_like_the_reference(lib):
    """The reference's reduction is CUDA max() = fmax (cuda_context.hpp:973-985): NaN elements are ignored,
    so one bad ray does not end workflow::converge_item for every other ray.  All-NaN and n == 0 give NaN."""
    ctx = lib.gfb_ctx_create(0)
    rng = np.random.default_rng(3)
    a = rng.normal(size=100000)
    a[::977] = np.nan
    out = ctypes.c_double(0)
    assert lib.gfb_buffer(ctx, 7001, a.nbytes, a.ctypes.data_as(ctypes.c_void_p), None) == 0
    assert lib.gfb_max(ctx, 7001, a.size, ctypes.byref(out)) == 0
    assert out.value == np.nanmax(a)
    b = np.full(1000, np.nan)
    assert lib.gfb_buffer(ctx, 7002, b.nbytes, b.ctypes.data_as(ctypes.c_void_p), None) == 0
    assert lib.gfb_max(ctx, 7002, b.size, ctypes.byref(out)) == 0 and np.isnan(out.value)
    assert lib.gfb_max(ctx, 7002, 0, ctypes.byref(out)) == 0 and np.isnan(out.value)
    lib.gfb_ctx_destroy(ctx)


def test_converge_item_keeps_iterating_past_a_nan_ray(lib):
    """Ensemble Newton (the reference's converge_item loop on the maximum residual) with one ray that can
    only produce NaN: the other rays must still converge."""
    from graph_framework_b200.rays import RayTracer
    from graph_framework_b200 import workloads
    n = 256
    state = workloads.efit_ensemble(n, seed=21)
    state["w"][17] = np.nan
    tr = RayTracer("extra_ordinary_wave", "efit", n, 2.0e-5)
    tr.set_state(state)
    tr.init("kx", mode="ensemble")
    got = tr.get_state(residual=False)
    tr.close()
    good = np.arange(n) != 17
    clean = dict(state, w=np.where(good, state["w"], 700.0))
    tr = RayTracer("extra_ordinary_wave", "efit", n, 2.0e-5)
    tr.set_state(clean)
    tr.init("kx", mode="ensemble")
    ref = tr.get_state(residual=False)
    tr.close()
    assert np.isnan(got["kx"][17])
    assert np.max(np.abs(got["kx"][good] - ref["kx"][good])/np.abs(ref["kx"][good])) < 1.0e-12
    assert np.max(np.abs(ref["kx"][good] + 700.0)) > 1.0          # and the solve did move them


def test_deposit_matches_oracle(lib):
    """Deposition histogram kernel vs the numpy restatement of utilities/bin.py:53-106."""
    import torch
    from graph_framework_b200 import workloads, parallel
    from graph_framework_b200.rays import RayTracer
    from oracle import port
    n = 20000
    state = workloads.efit_ensemble(n, seed=9)
    tr = RayTracer("ordinary_wave", "efit", n, 2.0e-5)
    tr.set_state(state)
    tr.init("kx")
    tr.compile()
    tr.step(50)
    got = tr.get_state()
    rng = np.random.default_rng(4)
    w = rng.uniform(0.0, 1.0, n)
    wt = torch.from_numpy(w).cuda()
    bins = (16, 16, 32)
    lo, hi = (2.3, -0.3, -0.3), (2.55, 0.3, 0.3)
    hist = torch.zeros(bins, dtype=torch.float64, device="cuda")
    torch.cuda.synchronize()
    parallel.deposit(tr, wt, hist, lo, hi)
    ref = port.deposit(got["x"], got["y"], got["z"], w, lo, hi, bins)
    tr.close()
    assert ref.sum() > 0.5*w.sum()
    assert np.allclose(hist.cpu().numpy(), ref, rtol=1.0e-12, atol=1.0e-12)



ADD_ONE = b'''
struct add_one_k {
    static constexpr int NI = 1, NR = 1, NG = 0, NP = 1, NE = 1, TI = -1;
    static constexpr unsigned SMEM_BYTES = 0, STAGED_BYTES = 0;
    __device__ static constexpr bool group_staged(const int) { return false; }
    __device__ static constexpr unsigned group_offset(const int) { return 0; }
    __device__ static constexpr unsigned group_bytes(const int) { return 0; }
    __device__ static constexpr int group_slot(const int g) { return NP + g; }
    __device__ static constexpr int ev(const int) { return 0; }
    __device__ static __forceinline__ void load(double (&v)[2], const gfb_args &a, const unsigned long long i) { v[0] = a.ptr[0][i]; }
    __device__ static __forceinline__ void apply(double (&v)[2], const double (&r)[2]) { v[0] = r[0]; }
    __device__ static __forceinline__ void store(const double (&v)[2], const double (&r)[2], const gfb_args &a, const unsigned long long i) { a.ptr[0][i] = v[0]; }
    static constexpr int NH = 0;
    __device__ static __forceinline__ void invariants(const double (&v)[2], double (&h)[1]) {}
    __device__ static __forceinline__ void body(const double (&v)[2], const double (&h)[1], double (&r)[2], const double *(&tg)[1]) { r[0] = v[0] + 1.0; }
};
extern "C" __global__ void add_one(const __grid_constant__ gfb_args a) { gfb::generic_item<add_one_k> (a); }
'''


def test_raw_c_abi_with_adopted_torch_memory(lib):
    """The device layer by hand: compile, buffer, kernel, deferred fused launches, and
    gfb_buffer_import re-pointing an existing kernel at torch-owned memory."""
    import torch
    ctx = lib.gfb_ctx_create(0)
    names = (ctypes.c_char_p*1)(b"add_one")
    assert lib.gfb_compile(ctx, ADD_ONE, names, 1, None) == 0, lib.gfb_last_error()
    n = 5000
    host = np.arange(n, dtype=np.float64)
    assert lib.gfb_buffer(ctx, 7, host.nbytes, host.ctypes.data_as(ctypes.c_void_p), None) == 0
    k = ctypes.c_void_p()
    keys = (ctypes.c_uint64*1)(7)
    assert lib.gfb_kernel_create(ctx, b"add_one", keys, 1, n, 128, 0, 0, 1, ctypes.byref(k)) == 0
    before = lib.gfb_launch_count(ctx)
    for _ in range(5):
        assert lib.gfb_kernel_run(k) == 0                   # deferred ...
    out = np.empty(n)
    assert lib.gfb_copy_d2h(ctx, 7, out.ctypes.data_as(ctypes.c_void_p), 0) == 0       # ... flushed here as one launch
    assert lib.gfb_launch_count(ctx) - before == 1
    assert np.array_equal(out, host + 5.0)
    t = torch.full((n,), 100.0, dtype=torch.float64, device="cuda")
    torch.cuda.synchronize()
    assert lib.gfb_buffer_import(ctx, 7, ctypes.c_void_p(t.data_ptr()), t.numel()*8) == 0
    assert lib.gfb_kernel_launch(k, 3) == 0
    assert lib.gfb_wait(ctx) == 0
    assert torch.equal(t.cpu(), torch.full((n,), 103.0, dtype=torch.float64))
    lib.gfb_ctx_destroy(ctx)


def test_index_kernels_and_pic_step(g):
    """index_1D/index_2D kernels (piecewise_test.cpp:834-925) and the three work items of
    graph_pic/xpic.cpp:106-131 with arrays of DIFFERENT lengths in one kernel: a field kernel of
    num_grid threads that walks num_particles particles (loop of batches, fused into one launch),
    then a particle push of num_particles threads that gathers from the num_grid field."""
    particles, grid_n, batch = 4096, 128, 8
    rng = np.random.default_rng(9)
    xs, vs = rng.normal(0.0, 0.3, particles), rng.normal(0.0, 0.25, particles)
    scale, offset, dt = 2.0/(grid_n - 1), -1.0, 1.0e-3
    x, v = g.variable(particles, "x", xs), g.variable(particles, "v", vs)
    e, dens = g.variable(grid_n, "e", np.ones(grid_n)), g.variable(grid_n, "n", np.ones(grid_n))
    pos = g.variable(grid_n, "xi", scale*np.arange(grid_n) + offset)
    idx = g.variable(grid_n, "i", np.full(grid_n, 5.0))

    def density(d):
        return g.exp(d*d/-0.01)

    next_i, next_e, next_n = idx, e, dens
    for _ in range(batch):
        d = g.index_1D(x, next_i, 1.0, 0.0) - pos
        nd = density(d)
        next_i = next_i + 1.0
        next_e = next_e + (-1.0/nd)*nd.df(d)
        next_n = next_n + nd
    zero = g.constant(0.0)
    g.add_item([idx, e, dens], [], [(zero, idx), (zero, e), (zero, dens)], "Index_reset", grid_n)
    g.add_item([e, dens, pos, idx, x], [], [(next_e, e), (next_i, idx), (next_n, dens)], "Compute_efield", grid_n)
    g.compile()
    g.run()                                   # reset + first batch
    lib_runs = particles//batch - 1

    def gather(arr, arg, s, o):
        return arr[np.clip((arg - o)/s, 0, arr.size - 1).astype(int)]

    grid = scale*np.arange(grid_n) + offset
    d = xs[:batch, None] - grid[None, :]
    assert np.allclose(g.copy_to_host(dens, grid_n), np.exp(d*d/-0.01).sum(axis=0), rtol=1e-13, atol=1e-300)
    assert np.array_equal(g.copy_to_host(idx, grid_n), np.full(grid_n, float(batch)))
    src = g.source()
    assert "group_slot" in src and src.count("tg[0][") >= batch


def test_index_2d_and_self_indexed_kernel(g):
    """index_2D on a variable grid, and a kernel that gathers from the array it rewrites: steps of
    such an item must not be fused (every run sees the array as the previous run left it)."""
    n = 256
    rng = np.random.default_rng(2)
    table = rng.normal(size=(8, 16))
    xs, ys = rng.uniform(-1.0, 9.0, n), rng.uniform(-1.0, 17.0, n)
    t = g.variable(table.size, "t", table.ravel())
    x, y = g.variable(n, "x", xs), g.variable(n, "y", ys)
    out = g.index_2D(t, 16, x, 1.0, 0.0, y, 1.0, 0.0)*2.0
    ring = g.variable(n, "ring", np.arange(n, dtype=float))
    left = g.variable(n, "left", (np.arange(n) + n - 1.0) % n)
    shifted = g.index_1D(ring, left, 1.0, 0.0)                       # ring[i] <- ring[i - 1]
    g.add_item([t, x, y, ring, left], [out], [(shifted, ring)], "gather", n)
    g.compile()
    for _ in range(5):
        g.run()
    ix, iy = np.clip(xs, 0, 7).astype(int), np.clip(ys, 0, 15).astype(int)
    assert np.array_equal(g.copy_to_host(out, n), 2.0*table[ix, iy])
    assert np.array_equal(g.copy_to_host(ring, n), np.roll(np.arange(n, dtype=float), 5))
