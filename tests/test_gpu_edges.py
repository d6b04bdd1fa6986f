"""Edge cases of the device layer and the ray path (GPU): ragged and tiny ensembles, an empty one,
error reporting, repeated compilation, step-count limits."""
import ctypes

import numpy as np
import pytest

from conftest import golden, rel_dev

pytestmark = pytest.mark.gpu
ORDER = ("t", "w", "x", "y", "z", "kx", "ky", "kz")


def unpack(a):
    return {k: np.array(a[i]) for i, k in enumerate(ORDER)}


@pytest.mark.parametrize("n", [1, 2, 31, 127, 129, 1000])
def test_ragged_sizes_match_full_run(lib, n):
    """Any ensemble size (not a multiple of the 128-thread block) gives the same per-ray results."""
    from graph_framework_b200.rays import RayTracer
    g = golden("ref_trace_ordinary_wave_efit_rk4")
    rec = g["per_step"]
    base = unpack(rec[0][:8])
    idx = np.arange(n) % base["w"].size
    start = {k: v[idx] for k, v in base.items()}
    tr = RayTracer("ordinary_wave", "efit", n, float(g["dt"]))
    tr.set_state(start)
    tr.init("")
    tr.compile()
    tr.step(5)
    got = tr.get_state()
    tr.close()
    for i, k in enumerate(ORDER):
        assert rel_dev(got[k], rec[5][i][idx]) < 1.0e-11, (n, k)


def test_empty_ensemble_is_a_no_op(lib):
    from graph_framework_b200.rays import RayTracer
    tr = RayTracer("simple", "slab", 0, 0.1)
    tr.set_state({k: np.zeros(0) for k in ORDER})
    tr.init("")
    tr.compile()
    tr.step(3)
    got = tr.get_state()
    tr.close()
    assert all(got[k].size == 0 for k in ORDER)


def test_step_counts_beyond_the_fusion_limit(lib):
    """More steps than one launch may fuse (default 1024) are split into several launches."""
    from graph_framework_b200.rays import RayTracer
    n, dt = 16, 1.0e-3
    s = {k: np.zeros(n) for k in ORDER}
    s["w"][:] = 1.0
    s["kx"][:] = 0.6
    s["ky"][:] = 0.5
    s["kz"][:] = 0.3
    tr = RayTracer("simple", "slab", n, dt)
    tr.set_state(s)
    tr.init("kx")
    tr.compile()
    before = tr.launch_count()
    tr.step(2500)
    got = tr.get_state()
    launches = tr.launch_count() - before
    tr.close()
    assert launches == 3                                   # 1024 + 1024 + 452
    assert np.allclose(got["t"], 2500*dt, rtol=1e-12)
    # D = k^2/w^2 - 1: dx/dt = k/w (|dx/dt| = 1), so |x| = t
    r = np.sqrt(got["x"]**2 + got["y"]**2 + got["z"]**2)
    assert np.allclose(r, 2500*dt, rtol=1e-10)


def test_compile_error_is_reported_not_swallowed(lib):
    """The reference prints the NVRTC log and carries on (cuda_context.hpp:256-267); here the call fails."""
    ctx = lib.gfb_ctx_create(0)
    assert ctx
    names = (ctypes.c_char_p*1)(b"broken")
    rc = lib.gfb_compile(ctx, b'extern "C" __global__ void broken(const gfb_args a) { this is not CUDA; }', names, 1, None)
    assert rc != 0
    assert b"error" in lib.gfb_last_error().lower()
    k = ctypes.c_void_p()
    keys = (ctypes.c_uint64*1)(1)
    assert lib.gfb_kernel_create(ctx, b"broken", keys, 0, 4, 128, 0, 0, 0, ctypes.byref(k)) != 0
    assert b"nothing compiled" in lib.gfb_last_error()
    assert lib.gfb_copy_d2h(ctx, 12345, ctypes.c_void_p(), 8) != 0
    assert b"unknown key" in lib.gfb_last_error()
    lib.gfb_ctx_destroy(ctx)


def test_device_info_is_a_b200(lib):
    ctx = lib.gfb_ctx_create(0)
    name = ctypes.create_string_buffer(128)
    sms, major, minor = ctypes.c_int(), ctypes.c_int(), ctypes.c_int()
    assert lib.gfb_ctx_device_info(ctx, name, 128, ctypes.byref(sms), ctypes.byref(major), ctypes.byref(minor)) == 0
    assert major.value == 10 and sms.value >= 100
    peak, ms = ctypes.c_double(), ctypes.c_float()
    assert lib.gfb_measure_fp64_peak(ctx, ctypes.byref(peak), ctypes.byref(ms)) == 0
    assert 20.0 < peak.value < 45.0                        # nominal 37.2 TFLOP/s
    lib.gfb_ctx_destroy(ctx)


def test_two_solvers_on_one_device_do_not_interfere(lib):
    """Several managers per device are allowed (workflow.hpp:233-238); buffers are per context."""
    from graph_framework_b200.rays import RayTracer
    g = golden("ref_trace_extra_ordinary_wave_efit_rk4")
    go = golden("ref_trace_ordinary_wave_efit_rk4")
    n = g["per_step"].shape[2]
    a = RayTracer("extra_ordinary_wave", "efit", n, float(g["dt"]))
    b = RayTracer("ordinary_wave", "efit", n, float(go["dt"]))
    for tr, rec in ((a, g["per_step"]), (b, go["per_step"])):
        tr.set_state(unpack(rec[0][:8]))
        tr.init("")
        tr.compile()
    a.step(3)
    b.step(5)
    a.step(2)
    ga, gb = a.get_state(), b.get_state()
    a.close()
    b.close()
    for i, k in enumerate(ORDER):
        assert rel_dev(ga[k], g["per_step"][5][i]) < 1.0e-11, ("x-mode", k)
        assert rel_dev(gb[k], go["per_step"][5][i]) < 1.0e-11, ("o-mode", k)


def test_xrays_driver_efit_example(lib, tmp_path):
    """The reference's efit_example.sh command line (graph_driver/efit_example.sh) at reduced size
    through graph_framework_b200.xrays: per-shard result files with (time, num_rays) variables."""
    from graph_framework_b200 import xrays
    from graph_framework_b200.tools.gfbt import read_gfbt
    prefix = str(tmp_path / "result")
    rc = xrays.main(["--dispersion=ordinary_wave", "--endtime=0.02", "--equilibrium=efit", "--init_kx",
                     "--init_kx_mean=-700.0", "--init_ky_dist=normal", "--init_ky_mean=-100.0", "--init_ky_sigma=10.0",
                     "--init_kz_dist=normal", "--init_kz_mean=0.0", "--init_kz_sigma=10.0", "--init_w_dist=normal",
                     "--init_w_mean=700", "--init_w_sigma=10.0", "--init_x_mean=2.5", "--init_y_dist=normal",
                     "--init_y_mean=0.0", "--init_y_sigma=0.05", "--init_z_dist=normal", "--init_z_mean=0.0",
                     "--init_z_sigma=0.05", "--num_rays=5000", "--num_times=1000", "--solver=rk4", "--sub_steps=100",
                     "--use_cyl_xy", "--seed", "--devices=1", "--output=" + prefix])
    assert rc == 0
    out = read_gfbt(prefix + "0.gfbt")
    assert out["x"].shape == (11, 5000)                     # num_times/sub_steps + 1: record 0 is the initial state (xrays.cpp:246-258)
    assert np.allclose(out["t"][:, 0], 2.0e-5*100*np.arange(0, 11), rtol=1e-12)
    assert np.allclose(np.sqrt(out["x"][0]**2 + out["y"][0]**2), 2.5, rtol=1e-14) and np.all(out["residual"][0] == 0.0)
    assert np.isfinite(out["kx"]).all() and np.max(out["residual"][-1]) < 1.0e-18
    r = np.sqrt(out["x"]**2 + out["y"]**2)
    assert np.all(r[-1] < r[0])                    # launched inward from R = 2.5


def test_ray_binning_is_invisible_and_exact(lib):
    """Rays of tabulated equilibria are kept sorted by table cell while stepping (automatic (R, Z)
    binning for EFIT, gfb_rays_set_binning for a 1-D grid).  Each ray's arithmetic does not depend on
    the slot it occupies, so state, residual and trajectory records must be BIT-IDENTICAL to the run
    without binning, in the caller's order, also across re-sorts."""
    from graph_framework_b200.rays import RayTracer
    from graph_framework_b200 import workloads
    n = 20000
    state = workloads.efit_ensemble(n, seed=12)
    runs = []
    for variant in ("bin_rays=0", "bin_rays=150", "1d"):
        tr = RayTracer("extra_ordinary_wave", "efit", n, 2.0e-5, options=None if variant == "1d" else variant)
        tr.set_state(state)
        tr.init("kx")
        tr.compile()
        if variant == "1d":
            tr.set_binning("y", -0.3, 0.3, 64, rebin_every=150)
        tr.step(100)
        mid = tr.get_state()
        rec = tr.trace(3, 100)
        tr.step(50)
        tr.step(50)
        end = tr.get_state()
        runs.append((mid, rec.copy(), end))
        tr.close()
    for other in runs[1:]:
        for a, b in zip(runs[0], other):
            if isinstance(a, dict):
                for k in a:
                    assert np.array_equal(a[k], b[k]), k
            else:
                assert np.array_equal(a, b)


def test_bin_rays_sorts_by_cell_and_unbin_restores(lib):
    """The device-layer calls by hand: after gfb_bin_rays every listed array is the same permutation
    of its input, sorted by the cell of the key array; gfb_unbin_rays restores the input."""
    n, cells = 100003, 37
    rng = np.random.default_rng(21)
    key = rng.uniform(-0.2, 1.2, n)
    key[:5] = [np.nan, -5.0, 7.0, 0.0, 1.0]
    other = np.arange(n, dtype=np.float64)
    ctx = lib.gfb_ctx_create(0)
    for k, a in ((11, key), (12, other)):
        assert lib.gfb_buffer(ctx, k, a.nbytes, a.ctypes.data_as(ctypes.c_void_p), None) == 0
    keys = (ctypes.c_uint64*2)(11, 12)
    assert lib.gfb_bin_rays(ctx, 11, 0.0, 1.0, cells, keys, 2, n) == 0, lib.gfb_last_error()
    assert lib.gfb_is_binned(ctx) == 1
    k2, o2 = np.empty(n), np.empty(n)
    assert lib.gfb_copy_d2h(ctx, 11, k2.ctypes.data_as(ctypes.c_void_p), 0) == 0
    assert lib.gfb_copy_d2h(ctx, 12, o2.ctypes.data_as(ctypes.c_void_p), 0) == 0
    perm = o2.astype(np.int64)
    assert np.array_equal(np.sort(perm), np.arange(n))
    assert np.array_equal(k2, key[perm], equal_nan=True)
    with np.errstate(invalid="ignore"):
        cell = np.clip(np.nan_to_num(k2*cells, nan=0.0), 0, cells - 1).astype(int)
    assert np.all(np.diff(cell) >= 0)
    assert lib.gfb_unbin_rays(ctx, keys, 2, n) == 0 and lib.gfb_is_binned(ctx) == 0
    assert lib.gfb_copy_d2h(ctx, 11, k2.ctypes.data_as(ctypes.c_void_p), 0) == 0
    assert lib.gfb_copy_d2h(ctx, 12, o2.ctypes.data_as(ctypes.c_void_p), 0) == 0
    assert np.array_equal(k2, key, equal_nan=True) and np.array_equal(o2, other)
    lib.gfb_ctx_destroy(ctx)


def test_boris_binning_is_invisible_and_exact(lib):
    """gfb_boris_set_binning ((R, Z) cells of the field tables, gfb_bin_rays_rz): the pushed particles
    are bit-identical to the unbinned run and come back in the caller's order."""
    from graph_framework_b200.rays import BorisPusher
    from graph_framework_b200 import workloads
    n = 50000
    start = workloads.boris_ensemble(n, seed=4)
    out = []
    for binning in (False, True):
        b = BorisPusher("efit", n, dt=0.5, options="bin_rays=0")
        b.set_state(*start)
        b.compile()
        if binning:
            b.set_binning((0.84, 0.84 + 64*0.0265625, 64), (-1.6, 1.6, 64), rebin_every=30)
        b.step(70)
        mid = b.get_state()
        b.step(50)
        out.append((mid, b.get_state()))
        b.close()
    for a, c in zip(out[0], out[1]):
        for k in a:
            assert np.array_equal(a[k], c[k]), k


def test_two_tracers_in_two_host_threads(lib):
    """The reference's model is one device context per host thread (xrays.cpp:419-527); the xrays
    driver here uses the same.  Two tracers built, compiled and stepped concurrently from two
    threads (same device) must give exactly what they give one after the other."""
    import threading
    from graph_framework_b200.rays import RayTracer
    from graph_framework_b200 import workloads

    def run(disp, seed, out, key):
        n = 30000
        tr = RayTracer(disp, "efit", n, 2.0e-5)
        tr.set_state(workloads.efit_ensemble(n, seed=seed))
        tr.init("kx")
        tr.compile()
        tr.step(150)
        rec = tr.trace(2, 50)
        out[key] = (tr.get_state(), rec.copy())
        tr.close()

    serial, threaded = {}, {}
    run("extra_ordinary_wave", 1, serial, "a")
    run("ordinary_wave", 2, serial, "b")
    threads = [threading.Thread(target=run, args=("extra_ordinary_wave", 1, threaded, "a")),
               threading.Thread(target=run, args=("ordinary_wave", 2, threaded, "b"))]
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    for key in ("a", "b"):
        assert key in threaded
        for k in serial[key][0]:
            assert np.array_equal(serial[key][0][k], threaded[key][0][k]), (key, k)
        assert np.array_equal(serial[key][1], threaded[key][1]), key


def test_rebinning_composes_permutations(lib):
    """Sorting rays that are already sorted (after they moved) folds the new permutation into the one
    in force: gfb_unbin_rays still restores the caller's order in one pass, and gfb_bin_disorder
    reports the decay that triggers the re-sort."""
    n, cells = 50021, 29
    rng = np.random.default_rng(33)
    key = rng.uniform(0.0, 1.0, n)
    tag = np.arange(n, dtype=np.float64)
    ctx = lib.gfb_ctx_create(0)
    for k, a in ((21, key), (22, tag)):
        assert lib.gfb_buffer(ctx, k, a.nbytes, a.ctypes.data_as(ctypes.c_void_p), None) == 0
    keys = (ctypes.c_uint64*2)(21, 22)
    sort_key = (ctypes.c_uint64*1)(21)
    lo, hi = np.array([0.0, 0.0]), np.array([1.0, 1.0])
    grid = (ctypes.c_uint*2)(cells, 0)

    def disorder():
        f = ctypes.c_double(0.0)
        assert lib.gfb_bin_disorder(ctx, sort_key, 1, lo.ctypes.data_as(ctypes.POINTER(ctypes.c_double)),
                                    hi.ctypes.data_as(ctypes.POINTER(ctypes.c_double)), grid, n, ctypes.byref(f)) == 0
        return f.value

    def fetch(k):
        out = np.empty(n)
        assert lib.gfb_copy_d2h(ctx, k, out.ctypes.data_as(ctypes.c_void_p), 0) == 0
        return out

    assert disorder() > 0.9                                            # random order: almost every neighbour differs
    assert lib.gfb_bin_rays(ctx, 21, 0.0, 1.0, cells, keys, 2, n) == 0
    assert disorder() < 2.0*cells/n                                    # sorted: one boundary per occupied cell
    first = fetch(22).astype(np.int64)
    # the rays "move": new key values, written in the CURRENT device order
    moved = np.clip(fetch(21) + rng.normal(0.0, 0.05, n), 0.0, 1.0)
    assert lib.gfb_copy_h2d(ctx, 21, moved.ctypes.data_as(ctypes.c_void_p), 0) == 0
    assert disorder() > 1.0/32.0
    assert lib.gfb_bin_rays(ctx, 21, 0.0, 1.0, cells, keys, 2, n) == 0      # re-sort, permutations composed
    total = fetch(22).astype(np.int64)
    assert np.array_equal(np.sort(total), np.arange(n))
    now = fetch(21)
    assert np.all(np.diff(np.clip(now*cells, 0, cells - 1).astype(int)) >= 0)
    # ray r = first[i] carried moved[i]; after the re-sort slot j holds ray total[j]
    carried = np.empty(n)
    carried[first] = moved
    assert np.array_equal(now, carried[total])
    host = np.empty(n)
    assert lib.gfb_copy_rays_d2h(ctx, 21, host.ctypes.data_as(ctypes.c_void_p), n) == 0     # caller's order, device untouched
    assert np.array_equal(host, carried) and lib.gfb_is_binned(ctx) == 1
    assert lib.gfb_unbin_rays(ctx, keys, 2, n) == 0
    assert np.array_equal(fetch(22), tag) and np.array_equal(fetch(21), carried)
    lib.gfb_ctx_destroy(ctx)
