"""GPU parity tests: the CUDA path (through the C ABI) against the reference oracle.

Golden vectors in tests/golden/ref_*.npz were produced by the UNMODIFIED reference code
(oracle/_ref/ref_driver, see oracle/make_golden.py).  Tolerances:
  * right-hand side and one step from identical pre-step state: 1e-12 relative (north star),
    with an absolute floor of 1e-9 x the ensemble scale for components that are ~0;
  * trajectories: stated per test.
cold_plasma + EFIT: the reference's symbolic dD/dz is defective (a reducer rule,
tests/test_oracle.py::test_reference_reducer_defect_is_pinned).  Its effect has a closed form
(dispersion::cold_plasma::reference_defect) and is reproduced by default, so this case is compared with
the REFERENCE like every other; with the option reference_defects=0 the true derivative is used and
compared with the independent numpy restatement (oracle/port.py).  For VMEC no closed form is known: there
dk/dt of cold plasma is compared with the reference's own D through finite differences elsewhere.
"""
import numpy as np
import pytest

from conftest import golden, rel_dev, assert_rhs_close

pytestmark = pytest.mark.gpu

ORDER = ("t", "w", "x", "y", "z", "kx", "ky", "kz")
RHS = ("dxdt", "dydt", "dzdt", "dkxdt", "dkydt", "dkzdt", "D")


def unpack(a):
    return {k: np.array(a[i]) for i, k in enumerate(ORDER)}


RHS_CASES = [("ordinary_wave", "efit"), ("extra_ordinary_wave", "efit"), ("cold_plasma", "efit"),
             ("cold_plasma", "slab"), ("cold_plasma", "slab_density"), ("ordinary_wave", "slab_density"),
             ("bohm_gross", "no_magnetic_field"), ("simple", "slab"), ("cold_plasma", "gaussian_density"),
             ("ordinary_wave", "vmec"), ("cold_plasma", "vmec")]
REFERENCE_DEFECT = {("cold_plasma", "vmec"): ("dkxdt", "dkydt", "dkzdt")}


@pytest.mark.parametrize("disp,eq", RHS_CASES)
def test_rhs_matches_reference(lib, disp, eq):
    """jit_test.cpp:358-425 analogue: every ray-equation component as a kernel vs the reference."""
    from graph_framework_b200.rays import RayTracer
    g = golden("ref_rhs_%s_%s" % (disp, eq))
    state = unpack(g["state"])
    n = state["w"].size
    tr = RayTracer(disp, eq, n, 1.0e-3)
    tr.set_state(state)
    got = tr.rhs()
    tr.close()
    for i, k in enumerate(RHS):
        if k in REFERENCE_DEFECT.get((disp, eq), ()):
            continue        # reference defect, see module docstring
        assert_rhs_close(got[k], g["rhs"][i], (disp, eq, k))


@pytest.mark.parametrize("defects", [True, False])
def test_rhs_cold_plasma_efit_matches_port_in_both_modes(lib, efit_tables, defects):
    """reference_defects=1 (default): the reference's effective dD/dz; =0: the true derivative.  Both
    against the numpy restatement, whose derivatives are complex-step."""
    from graph_framework_b200.rays import RayTracer
    from oracle import port
    g = golden("ref_rhs_cold_plasma_efit")
    state = unpack(g["state"])
    tr = RayTracer("cold_plasma", "efit", state["w"].size, 1.0e-3, options="reference_defects=%d" % defects)
    tr.set_state(state)
    got = tr.rhs()
    tr.close()
    with np.errstate(all="ignore"):
        ref = port.rhs("cold_plasma", port.Efit(efit_tables), state, reference_defects=defects)
    for k in RHS:
        assert_rhs_close(got[k], ref[k], ("cold_plasma", "efit", k, "vs port", defects))
    if not defects:
        assert np.median(np.abs(got["dkzdt"] - g["rhs"][5])/np.abs(g["rhs"][5])) > 0.1


#  (dispersion, equilibrium, solver, golden tag).  cold_plasma + EFIT: "efit_interior" starts inside the
#  plasma; "efit" are the efit_example rays, which start at R = 2.5 in VACUUM where cold-plasma D is doubly
#  degenerate (D ~ (1 - n^2)^2): there the reference's dkz/dt is its reducer defect times (n^2 - 1) ~ 3e-9, a
#  difference of O(1) numbers known to ~1e-7 relative to ANY evaluation order, so kz of that case is held
#  to 1e-6 per step and everything else to 1e-12.
TRACE_CASES = [("extra_ordinary_wave", "efit", "rk4", "efit"), ("ordinary_wave", "efit", "rk4", "efit"),
               ("cold_plasma", "efit", "rk4", "efit_interior"), ("cold_plasma", "efit", "rk4", "efit"),
               ("cold_plasma", "slab_density", "rk4", "slab_density"), ("ordinary_wave", "slab_density", "rk4", "slab_density"),
               ("cold_plasma", "slab", "rk2", "slab"), ("simple", "slab", "rk4", "slab"),
               # the symplectic split (solver.hpp:1017-1130) on the two separable Hamiltonians
               ("bohm_gross", "no_magnetic_field", "split_simplextic", "no_magnetic_field"),
               ("light_wave", "no_magnetic_field", "split_simplextic", "no_magnetic_field")]
ILL_CONDITIONED = {("cold_plasma", "efit"): {"kz": 1.0e-6}}
#  The residual is D^2 at a Newton root = the square of D's rounding noise; for cold plasma + EFIT that
#  noise is ~1e-11 (folded spline coefficients up to 4e7, conftest.assert_rhs_close), elsewhere < 1e-14.
#  bohm_gross / light_wave: D = w_pe^2 + ... - w^2 is a difference of terms ~1.2e6, one ulp of which is 2.3e-10.
RESIDUAL_FLOOR = {("cold_plasma", "efit"): 1.0e-10, ("cold_plasma", "efit_interior"): 1.0e-10,
                  ("bohm_gross", "no_magnetic_field"): 2.0e-9, ("light_wave", "no_magnetic_field"): 2.0e-9}


@pytest.mark.parametrize("disp,eq,solver,tag", TRACE_CASES)
@pytest.mark.parametrize("mode", ["per_ray", "ensemble"])
def test_newton_init_matches_reference(lib, disp, eq, solver, tag, mode):
    """dispersion_test.cpp:25-64 analogue: Newton solve for kx; compare the converged wavenumber."""
    from graph_framework_b200.rays import RayTracer
    g = golden("ref_trace_%s_%s_%s" % (disp, tag, solver))
    state = unpack(g["state"])
    n = state["w"].size
    tr = RayTracer(disp, eq, n, float(g["dt"]), solver=solver)
    tr.set_state(state)
    tr.init("kx", mode=mode)
    got = tr.get_state(residual=False)
    tr.close()
    ref = g["per_step"][0]
    # Newton stops on a residual of 1e-30 in D^2: converged roots agree to ~1e-14 relative.  Cold plasma +
    # EFIT: D is flat near its root in vacuum (double root) and noisy inside (RESIDUAL_FLOOR): the root is
    # defined to |D noise|/|dD/dkx|: observed 8e-12 inside the plasma, 7e-9 on the vacuum double root.
    tol = {("cold_plasma", "efit_interior"): 1.0e-10, ("cold_plasma", "efit"): 1.0e-7}.get((disp, tag), 1.0e-12)
    assert rel_dev(got["kx"], ref[5]) < tol, rel_dev(got["kx"], ref[5])


@pytest.mark.parametrize("disp,eq,solver,tag", TRACE_CASES)
def test_one_step_from_identical_state(lib, disp, eq, solver, tag):
    """Per-step parity: copy the reference's pre-step state to the device, one step, compare
    all outputs and the residual (D^2 at the pre-step state, solver.hpp:316-319)."""
    from graph_framework_b200.rays import RayTracer
    g = golden("ref_trace_%s_%s_%s" % (disp, tag, solver))
    loose = ILL_CONDITIONED.get((disp, tag), {})
    rec = g["per_step"]
    n = rec.shape[2]
    tr = RayTracer(disp, eq, n, float(g["dt"]), solver=solver)
    tr.set_state(unpack(rec[0][:8]))
    tr.init("")
    tr.compile()
    for step in range(rec.shape[0] - 1):
        tr.put_state(unpack(rec[step][:8]))
        tr.step(1)
        got = tr.get_state()
        for i, k in enumerate(ORDER):
            assert rel_dev(got[k], rec[step + 1][i]) < loose.get(k, 1.0e-12), (step, k, rel_dev(got[k], rec[step + 1][i]))
        d_ref = np.sqrt(rec[step + 1][8])
        assert np.max(np.abs(np.sqrt(got["residual"]) - d_ref)) <= 1.0e-12*np.max(d_ref) + RESIDUAL_FLOOR.get((disp, tag), 1.0e-14)
    tr.close()


@pytest.mark.parametrize("disp,eq,solver,tag", TRACE_CASES)
@pytest.mark.parametrize("graph_stages", [False, True])
def test_trajectory_matches_reference(lib, disp, eq, solver, tag, graph_stages):
    """200 fused steps (two launches of 100) from the reference's post-Newton state; stated
    end-of-trajectory tolerance 1e-9 relative (observed ~1e-12; errors grow along the ray).  The vacuum-start
    cold-plasma case is ill-conditioned in kz step after step (ILL_CONDITIONED) and is left to the per-step test."""
    from graph_framework_b200.rays import RayTracer
    if (disp, tag) in ILL_CONDITIONED:
        pytest.skip("kz of the reference itself is rounding-dominated along this trajectory; per-step test covers it")
    if graph_stages and not solver.startswith("rk"):
        pytest.skip("only rk2/rk4 have a skeleton and a graph-unrolled construction")
    g = golden("ref_trace_%s_%s_%s" % (disp, tag, solver))
    rec = g["long"]
    n = rec.shape[2]
    tr = RayTracer(disp, eq, n, float(g["dt"]), solver=solver + ("_graph" if graph_stages else ""))
    tr.set_state(unpack(rec[0][:8]))
    tr.init("")
    tr.compile()
    for block in (1, 2):
        tr.step(100)
        got = tr.get_state()
        for i, k in enumerate(ORDER):
            assert rel_dev(got[k], rec[block][i]) < 1.0e-9, (block, k, rel_dev(got[k], rec[block][i]))
    tr.close()


def test_adaptive_rk4_runs_the_reference_rule(lib):
    """solver::adaptive_rk4 (solver.hpp:882-1006) as written: before every step a Newton solve on (dt, lambda)
    of 1/dt + lambda D_next^2.  The rule is ill-posed -- the REFERENCE's own run of this case
    (tests/golden/ref_trace_cold_plasma_gaussian_density_adaptive_rk4.npz) leaves |dt| > 1e10 or NaN on most
    rays after ONE step -- so there is no numerical target to hold a second implementation to; checked here:
    the golden documents that, and this back end runs the same two work items per step with a per-ray dt."""
    from graph_framework_b200.rays import RayTracer
    g = golden("ref_trace_cold_plasma_gaussian_density_adaptive_rk4")
    ref_dt = g["records"][1][9]
    assert np.mean(~np.isfinite(ref_dt) | (np.abs(ref_dt) > 1.0e10)) > 0.7          # 24 of 32 rays; the other 8 have |dt| ~ 1e-2 .. 1e9
    state = unpack(g["state"])
    n = state["w"].size
    tr = RayTracer("cold_plasma", "gaussian_density", n, float(g["dt"]), solver="adaptive_rk4")
    tr.set_state(state)
    tr.init("kx")
    start = tr.get_state(residual=False)
    assert rel_dev(start["kx"], g["records"][0][5]) < 1.0e-12           # the Newton root before stepping does match
    tr.compile()
    before = tr.launch_count()
    tr.step(1)
    dt = tr.get_dt()
    got = tr.get_state()
    assert tr.launch_count() - before == 2                               # the (dt, lambda) Newton item, then the step
    tr.close()
    assert dt.shape == (n,) and np.any(dt != float(g["dt"]))             # every ray chose its own step
    moved = np.isfinite(dt)
    assert np.array_equal(np.isfinite(got["t"]), moved)
    assert np.allclose(got["t"][moved], dt[moved], rtol=1.0e-14, atol=0.0)      # t advanced by the solved dt


def test_fused_steps_equal_single_steps(lib):
    """Deferred-launch fusion must not change results: 37 single-step launches == one 37-step launch, bit for bit."""
    from graph_framework_b200.rays import RayTracer
    g = golden("ref_trace_extra_ordinary_wave_efit_rk4")
    start = unpack(g["per_step"][0][:8])
    n = start["w"].size
    out = []
    for fused in (1, 64):
        tr = RayTracer("extra_ordinary_wave", "efit", n, float(g["dt"]), options="fused_steps=%d bin_rays=0" % fused)      # binning adds its own launches
        tr.set_state(start)
        tr.init("")
        tr.compile()
        before = tr.launch_count()
        tr.step(37)
        out.append((tr.get_state(), tr.launch_count() - before))
        tr.close()
    assert out[0][1] == 37 and out[1][1] == 1
    for k in ORDER + ("residual",):
        assert np.array_equal(out[0][0][k], out[1][0][k]), k


def test_xrays_bench_case(lib):
    """The reference benchmark's own initial conditions (xrays_bench.cpp:62-79): Newton gives
    kx = -500.0000036; x, kx, t match the reference after 10 steps.  z/kz are
    excluded: downstream of the reference's dD/dz defect."""
    from graph_framework_b200.rays import RayTracer
    g = golden("ref_trace_xrays_bench")
    rec = g["records"]
    state = unpack(g["state"])
    tr = RayTracer("cold_plasma", "efit", state["w"].size, float(g["dt"]))
    tr.set_state(state)
    tr.init("kx")
    k0 = tr.get_state(residual=False)["kx"]
    assert rel_dev(k0, rec[0][5]) < 1.0e-13
    assert abs(k0[0] + 500.0000036) < 1.0e-6
    tr.compile()
    tr.step(10)
    got = tr.get_state()
    tr.close()
    for k in ("t", "x", "kx"):
        i = ORDER.index(k)
        assert rel_dev(got[k], rec[-1][i]) < 1.0e-9, (k, rel_dev(got[k], rec[-1][i]))


@pytest.mark.parametrize("defects", [True, False])
def test_cold_plasma_efit_trajectory_matches_port(lib, efit_tables, defects):
    """20 steps inside the plasma in both derivative modes against the numpy restatement (the default
    mode is also held to the reference itself by the TRACE_CASES tests above)."""
    from graph_framework_b200.rays import RayTracer
    from oracle import port
    g = golden("ref_trace_cold_plasma_efit_interior_rk4")
    start = unpack(g["per_step"][0][:8])
    n = start["w"].size
    dt = float(g["dt"])
    tr = RayTracer("cold_plasma", "efit", n, dt, options="reference_defects=%d" % defects)
    tr.set_state(start)
    tr.init("")
    tr.compile()
    tr.step(20)
    got = tr.get_state()
    tr.close()
    with np.errstate(all="ignore"):
        ref, res = port.trace("cold_plasma", port.Efit(efit_tables), start, dt, 20, reference_defects=defects)
    for k in ORDER:
        assert rel_dev(got[k], ref[k]) < 1.0e-10, (k, rel_dev(got[k], ref[k]))


def test_boris_push_matches_reference(lib):
    """xkorc step graph: on-axis field from the damped Newton search, then 50 pushes."""
    from graph_framework_b200.rays import BorisPusher
    g = golden("ref_korc_efit")
    start = g["start"]
    b = BorisPusher("efit", start.shape[1])
    info = b.info()
    assert abs(info["b0"] - float(g["b0"])) < 1.0e-12*abs(float(g["b0"]))
    assert abs(info["larmor_radius"] - float(g["larmor_radius"])) < 1.0e-12*abs(float(g["larmor_radius"]))
    b.set_state(*start)
    b.compile()
    b.step(50)
    got = b.get_state()
    b.close()
    for i, k in enumerate(BorisPusher.NAMES):
        assert rel_dev(got[k], g["end"][i]) < 1.0e-10, (k, rel_dev(got[k], g["end"][i]))


def test_million_ray_properties(lib):
    """BASELINE size (10^6 rays, X-mode, EFIT): size-independent properties instead of an oracle run:
    Newton drives D^2 below 1e-20 on every ray, 100 RK4 steps keep the dispersion relation
    satisfied (solver_test.cpp:28-60 analogue), w is untouched, t advances by exactly 100 dt,
    and a permutation of the rays permutes the results (rays are independent)."""
    from graph_framework_b200 import workloads
    from graph_framework_b200.rays import RayTracer
    n, dt = 1000000, 2.0e-5
    state = workloads.efit_ensemble(n, seed=11)
    tr = RayTracer("extra_ordinary_wave", "efit", n, dt)
    tr.set_state(state)
    tr.init("kx")
    tr.compile()
    tr.step(100)
    got = tr.get_state()
    tr.close()
    assert np.isfinite(got["x"]).all() and np.isfinite(got["kx"]).all()
    assert np.array_equal(got["w"], state["w"])
    assert np.max(np.abs(got["t"] - 100*dt)) < 1.0e-15
    assert np.max(got["residual"]) < 1.0e-20
    perm = np.random.default_rng(0).permutation(n)[:4096]
    sub = {k: state[k][perm] for k in ORDER}
    tr = RayTracer("extra_ordinary_wave", "efit", perm.size, dt)
    tr.set_state(sub)
    tr.init("kx")
    tr.compile()
    tr.step(100)
    got2 = tr.get_state()
    tr.close()
    for k in ORDER:
        assert np.array_equal(got2[k], got[k][perm]), k
    # ... and an oracle check at full size: 512 of the 10^6 rays against the independent numpy
    # restatement (Newton + 100 RK4 steps), 1e-9 relative at the end of the block.
    from oracle import port
    from graph_framework_b200.tools.gfbt import read_gfbt
    import os
    from conftest import GOLDEN
    sel = perm[:512]
    eq = port.Efit(read_gfbt(os.path.join(GOLDEN, "efit.gfbt")))
    with np.errstate(all="ignore"):
        s0 = port.newton("extra_ordinary_wave", eq, {k: state[k][sel] for k in ORDER}, "kx", per_ray=True)
        ref, res = port.trace("extra_ordinary_wave", eq, s0, dt, 100)
    for k in ORDER:
        assert rel_dev(got[k][sel], ref[k]) < 1.0e-9, (k, rel_dev(got[k][sel], ref[k]))


def test_vmec_cold_plasma_dkdt_matches_reference_finite_differences(lib):
    """cold_plasma + VMEC dk/dt against 4th-order central differences of the reference's OWN D (its symbolic
    dk/dt is defective there, 5-16 % off its own differences); stated 1e-7 = accuracy of the differences."""
    from graph_framework_b200.rays import RayTracer
    g = golden("ref_fd_cold_plasma_vmec")
    state = unpack(g["state"])
    tr = RayTracer("cold_plasma", "vmec", state["w"].size, 1.0e-4)
    tr.set_state(state)
    got = tr.rhs()
    tr.close()
    assert rel_dev(got["D"], g["D"]) < 1.0e-12
    for k, fd in (("dkxdt", "dDdx"), ("dkydt", "dDdy"), ("dkzdt", "dDdz")):
        ref = g[fd]/g["dDdw"]
        assert rel_dev(got[k], ref) < 1.0e-7, (k, rel_dev(got[k], ref))


def test_vmec_trajectory_matches_reference(lib):
    """VMEC O-mode against a trajectory of the reference itself (oracle/make_golden.py vmec_trace): Newton
    root (both modes), three single steps from the reference's pre-step states at 1e-12, and 20 steps --
    fused, binned by radial cell -- at a stated 1e-9 (observed < 1e-15)."""
    from graph_framework_b200.rays import RayTracer
    g = golden("ref_trace_ordinary_wave_vmec_rk4")
    rec = g["per_step"]
    n = rec.shape[2]
    dt = float(g["dt"])
    for mode in ("per_ray", "ensemble"):
        tr = RayTracer("ordinary_wave", "vmec", n, dt)
        tr.set_state(unpack(g["state"]))
        tr.init("kx", mode=mode)
        got = tr.get_state(residual=False)
        assert rel_dev(got["kx"], rec[0][5]) < 1.0e-12, (mode, rel_dev(got["kx"], rec[0][5]))
        tr.close()
    tr = RayTracer("ordinary_wave", "vmec", n, dt)
    tr.set_state(unpack(rec[0][:8]))
    tr.init("")
    tr.compile()
    for step in range(3):
        tr.put_state(unpack(rec[step][:8]))
        tr.step(1)
        got = tr.get_state()
        for i, k in enumerate(ORDER):
            assert rel_dev(got[k], rec[step + 1][i]) < 1.0e-12, (step, k, rel_dev(got[k], rec[step + 1][i]))
        assert np.max(np.abs(np.sqrt(got["residual"]) - np.sqrt(rec[step + 1][8]))) < 1.0e-13
    tr.put_state(unpack(rec[0][:8]))
    tr.step(20)
    got = tr.get_state()
    tr.close()
    for i, k in enumerate(ORDER):
        assert rel_dev(got[k], rec[20][i]) < 1.0e-9, (k, rel_dev(got[k], rec[20][i]))


def test_vmec_trajectory_properties(lib):
    """VMEC (flux coordinates, 86 Fourier modes): no reference test constructs it and a reference RK4
    run costs ~17 minutes of graph build + compile, so beyond the right-hand-side parity above the
    trajectory is checked through properties: Newton puts every ray on D = 0, 20 RK4 steps keep
    D^2 small, rk4 and the graph-unrolled construction agree, and halving dt changes the end point
    by ~dt^4."""
    from graph_framework_b200 import workloads
    from graph_framework_b200.rays import RayTracer
    n = 64
    state = workloads.vmec_states(n, seed=12)
    ends = {}
    for dt, steps in ((1.0e-4, 20), (0.5e-4, 40)):
        tr = RayTracer("ordinary_wave", "vmec", n, dt)
        tr.set_state(state)
        tr.init("kx")
        start = tr.get_state(residual=False)
        tr.compile()
        tr.step(steps)
        ends[dt] = tr.get_state()
        tr.close()
        assert np.isfinite(ends[dt]["x"]).all()
        assert np.max(ends[dt]["residual"]) < 1.0e-12          # D^2 drift of RK4 (Newton start: ~1e-31)
        assert np.max(np.abs(ends[dt]["x"] - start["x"])) > 0.0
    a, b = ends[1.0e-4], ends[0.5e-4]
    for k in ("x", "y", "z", "kx", "ky", "kz"):
        scale = max(np.max(np.abs(b[k])), 1.0e-300)
        assert np.max(np.abs(a[k] - b[k]))/scale < 1.0e-6, k


def test_trajectory_output_matches_blockwise_reads(lib, tmp_path):
    """write_step analogue: overlapped snapshots equal stop-and-copy reads, and the file round-trips."""
    from graph_framework_b200.rays import RayTracer
    from graph_framework_b200.tools.gfbt import write_trajectory, read_gfbt
    g = golden("ref_trace_extra_ordinary_wave_efit_rk4")
    rec = g["long"]
    start = unpack(rec[0][:8])
    n = start["w"].size
    tr = RayTracer("extra_ordinary_wave", "efit", n, float(g["dt"]))
    tr.set_state(start)
    tr.init("")
    tr.compile()
    out = tr.trace(4, 50)
    tr.put_state(start)
    for b in range(4):
        tr.step(50)
        st = tr.get_state()
        for i, k in enumerate(ORDER + ("residual",)):
            assert np.array_equal(out[b][i], st[k]), (b, k)
    tr.close()
    for block, ref_block in ((1, 1), (3, 2)):          # 100 and 200 steps vs the reference
        for i, k in enumerate(ORDER):
            assert rel_dev(out[block][i], rec[ref_block][i]) < 1.0e-9, (block, k)
    path = str(tmp_path / "rays.gfbt")
    write_trajectory(path, out)
    back = read_gfbt(path)
    assert back["x"].shape == (4, n) and np.array_equal(back["kz"], out[:, 7, :])


def test_pipelined_host_stepping_is_bit_identical(lib):
    """gfb_rays_step_host (chunked upload / stepping / read-back overlap) == put_state + step + get_state."""
    from graph_framework_b200.rays import RayTracer
    g = golden("ref_trace_extra_ordinary_wave_efit_rk4")
    base = unpack(g["per_step"][0][:8])
    # ragged sizes: 22400 rays (less than one wave of blocks: equal pieces) and 400032 rays (7.04 waves:
    # one-wave first piece, tail-wave last piece, the rest shared by the pieces between)
    for reps, chunk_counts in ((700, (1, 3, 8)), (12501, (3,))):
        start = {k: np.tile(v, reps) for k, v in base.items()}
        n = start["w"].size
        tr = RayTracer("extra_ordinary_wave", "efit", n, float(g["dt"]))
        tr.set_state(start)
        tr.init("")
        tr.compile()
        tr.step(25)
        ref = tr.get_state()
        for chunks in chunk_counts:
            out = {k: np.empty(n) for k in ORDER + ("residual",)}
            tr.step_host(25, start, out, chunks=chunks)
            for k in ORDER + ("residual",):
                assert np.array_equal(out[k], ref[k]), (reps, chunks, k)
        tr.close()


def test_host_stepping_after_device_stepping_sees_the_new_input(lib):
    """gfb_rays_step_host right after gfb_rays_step on binned rays: the uploads must be ordered after the
    compute stream's pending work (the step's stores and the scatter restoring the caller's order), or the
    chunk kernels would start from the previous state."""
    from graph_framework_b200.rays import RayTracer
    from graph_framework_b200 import workloads
    n = 400000
    a = workloads.efit_ensemble(n, seed=31)
    b = workloads.efit_ensemble(n, seed=32)
    tr = RayTracer("extra_ordinary_wave", "efit", n, 2.0e-5)
    tr.set_state(a)
    tr.init("kx")
    tr.compile()
    kx_a = tr.get_state(residual=False)["kx"]
    # reference result for input b (own Newton root needed: reuse a's kx as b's, any finite value will do)
    b["kx"] = kx_a.copy()
    tr.put_state(b)
    tr.step(25)
    ref = tr.get_state()
    for _ in range(3):
        tr.put_state(dict(a, kx=kx_a))
        tr.step(300)                                    # long enough to be running (and binned) when step_host is called
        out = {k: np.empty(n) for k in ORDER + ("residual",)}
        tr.step_host(25, b, out, chunks=4)
        for k in ORDER + ("residual",):
            assert np.array_equal(out[k], ref[k]), k
    tr.close()
