"""Emitted bodies + skeleton step logic executed on the CPU (g++) against the reference's golden
vectors, and NVRTC compilation of the same text for sm_100a.  No GPU needed.

The CPU harness (tests/cpu_harness) is test infrastructure: it compiles skeleton.cuh with five PTX
primitives stubbed and runs one "thread" per ray.  It exists so that every change to the front end
or emitter is checked against the reference before any GPU time is spent."""
import ctypes
import json
import os
import subprocess

import numpy as np
import pytest

from conftest import ROOT, golden, rel_dev, assert_rhs_close

BUILD = os.path.join(ROOT, "build")
EMIT = os.path.join(BUILD, "emit_case")


@pytest.fixture(scope="session")
def emit_tool(lib):
    os.makedirs(BUILD, exist_ok=True)
    src = os.path.join(ROOT, "tests", "cpp", "emit_case.cpp")
    deps = [src] + [os.path.join(ROOT, "graph_framework_b200", "csrc", "graph", f)
                    for f in os.listdir(os.path.join(ROOT, "graph_framework_b200", "csrc", "graph"))]
    if not os.path.exists(EMIT) or any(os.path.getmtime(d) > os.path.getmtime(EMIT) for d in deps):
        subprocess.run(["g++", "-std=c++20", "-O1", "-I" + os.path.join(ROOT, "include"), "-I/usr/local/cuda/include",
                        src, "-L" + os.path.join(ROOT, "graph_framework_b200"), "-lgfb200",
                        "-Wl,-rpath," + os.path.join(ROOT, "graph_framework_b200"), "-o", EMIT], check=True, cwd=ROOT)
    return EMIT


def emit(tool, disp, eq, kind, tag, dt=None, true_derivatives=False):
    cu = os.path.join(BUILD, "emit_%s.cu" % tag)
    tab = os.path.join(BUILD, "emit_%s.tab" % tag)
    env = dict(os.environ)
    env.pop("GFB_TRUE_DERIVATIVES", None)
    if true_derivatives:
        env["GFB_TRUE_DERIVATIVES"] = "1"       # dispersion::reference_defects() = false
    if dt is not None:
        env["GFB_DT"] = repr(float(dt))
    out = subprocess.run([tool, disp, eq, kind, cu, tab], check=True, capture_output=True, text=True, cwd=ROOT, env=env).stdout
    return cu, tab, json.loads(out)


def run_harness(cu, tab, kernel, arrays, n, steps, ni, no, tag, scalar=None):
    exe = os.path.join(BUILD, "harness_%s" % tag)
    deps = [cu, os.path.join(ROOT, "tests", "cpu_harness", "harness.cpp"),
            os.path.join(ROOT, "graph_framework_b200", "csrc", "skeleton.cuh"),
            os.path.join(ROOT, "graph_framework_b200", "csrc", "special.cuh")]
    if not os.path.exists(exe) or any(os.path.getmtime(d) > os.path.getmtime(exe) for d in deps):
        subprocess.run(["g++", "-std=c++17", "-O1", "-ffp-contract=off", '-DGFB_KERNEL_FILE="%s"' % cu,
                        "-DGFB_KERNEL_NAME=%s" % kernel, deps[1], "-o", exe], check=True)
    fin, fout = os.path.join(BUILD, tag + "_in.bin"), os.path.join(BUILD, tag + "_out.bin")
    np.ascontiguousarray(arrays, dtype=np.float64).tofile(fin)
    cmd = [exe, tab, fin, fout, str(n), str(steps), str(ni), str(no)]
    if scalar is not None:
        cmd.append(repr(scalar))
    subprocess.run(cmd, check=True)
    return np.fromfile(fout).reshape(ni + no, n)


RHS_CASES = [("ordinary_wave", "efit"), ("extra_ordinary_wave", "efit"), ("cold_plasma", "efit"),
             ("cold_plasma", "slab"), ("cold_plasma", "slab_density"), ("ordinary_wave", "slab_density"),
             ("bohm_gross", "no_magnetic_field"), ("simple", "slab"), ("ordinary_wave", "vmec"), ("cold_plasma", "vmec")]
#  cold_plasma in a field that depends on the coordinate: the reference's symbolic dD/dx_i is defective
#  (tests/test_oracle.py::test_reference_dkz_defect).  For EFIT the defect has a closed form and is
#  reproduced by default (dispersion::cold_plasma::reference_defect); for VMEC those components are
#  not compared with the reference.
REFERENCE_DEFECT = {("cold_plasma", "vmec"): ("dkxdt", "dkydt", "dkzdt")}


@pytest.mark.parametrize("disp,eq", RHS_CASES)
def test_emitted_rhs_matches_reference(emit_tool, disp, eq):
    g = golden("ref_rhs_%s_%s" % (disp, eq))
    n = g["state"].shape[1]
    tag = "rhs_%s_%s" % (disp, eq)
    cu, tab, info = emit(emit_tool, disp, eq, "rhs", tag)
    assert info["divides"] == 0 or info["reciprocals"] >= 0
    out = run_harness(cu, tab, "rhs_kernel", g["state"], n, 1, 8, 7, tag)[8:]
    for i, k in enumerate(("dxdt", "dydt", "dzdt", "dkxdt", "dkydt", "dkzdt", "D")):
        if k in REFERENCE_DEFECT.get((disp, eq), ()):
            continue
        if np.max(np.abs(g["rhs"][i])) == 0.0:
            assert np.max(np.abs(out[i])) == 0.0
            continue
        assert_rhs_close(out[i], g["rhs"][i], (disp, eq, k))


def test_emitted_rhs_true_derivatives_match_port(emit_tool):
    """reference_defects() = false: cold plasma + EFIT with the true dD/dz, against the numpy
    restatement (complex-step derivatives, independent of both symbolic differentiators)."""
    import sys
    sys.path.insert(0, ROOT)
    from oracle import port
    from graph_framework_b200.tools.gfbt import read_gfbt
    g = golden("ref_rhs_cold_plasma_efit")
    n = g["state"].shape[1]
    tag = "rhs_cold_plasma_efit_true"
    cu, tab, info = emit(emit_tool, "cold_plasma", "efit", "rhs", tag, true_derivatives=True)
    out = run_harness(cu, tab, "rhs_kernel", g["state"], n, 1, 8, 7, tag)[8:]
    eq = port.Efit(read_gfbt(os.path.join(ROOT, "tests", "golden", "efit.gfbt")))
    with np.errstate(all="ignore"):
        ref = port.rhs("cold_plasma", eq, dict(zip(port.ORDER, g["state"])), reference_defects=False)
        quirk = port.rhs("cold_plasma", eq, dict(zip(port.ORDER, g["state"])))
    for i, k in enumerate(("dxdt", "dydt", "dzdt", "dkxdt", "dkydt", "dkzdt", "D")):
        assert_rhs_close(out[i], ref[k], ("cold_plasma", "efit", k, "true derivatives vs port"))
    # the port's own restatement of the defect equals the reference's kernels
    assert_rhs_close(quirk["dkzdt"], g["rhs"][5], "port with the reference's defect vs reference")
    assert np.median(np.abs(ref["dkzdt"] - g["rhs"][5])/np.abs(g["rhs"][5])) > 0.1     # and the defect is not small


#  (dispersion, equilibrium, solver, golden tag).  cold_plasma + EFIT: "efit_interior" starts inside the
#  plasma; "efit" are the efit_example rays, which start at R = 2.5 in VACUUM where cold-plasma D is
#  doubly degenerate (D ~ (1 - n^2)^2): there the reference's dkz/dt is its reducer defect times
#  (n^2 - 1) ~ 3e-9, a difference of O(1) numbers known to ~1e-7 relative to ANY evaluation order, so
#  kz of that case is held to 1e-6 per step and everything else to 1e-12.
TRACE_CASES = [("extra_ordinary_wave", "efit", "rk4", "efit"), ("ordinary_wave", "efit", "rk4", "efit"),
               ("cold_plasma", "efit", "rk4", "efit_interior"), ("cold_plasma", "efit", "rk4", "efit"),
               ("cold_plasma", "slab_density", "rk4", "slab_density"), ("cold_plasma", "slab", "rk2", "slab")]
ILL_CONDITIONED = {("cold_plasma", "efit"): {7: 1.0e-6}}
#  The residual is D^2 at a Newton root, i.e. the square of D's rounding noise.  For cold plasma + EFIT
#  that noise is ~1e-11 (psi ~ 0.3 is a sum of folded-spline terms up to 4e7, conftest.assert_rhs_close),
#  so |D| is compared with that absolute floor; elsewhere the floor is 1e-14.
RESIDUAL_FLOOR = {("cold_plasma", "efit"): 1.0e-10, ("cold_plasma", "efit_interior"): 1.0e-10}


@pytest.mark.parametrize("disp,eq,solver,tag", TRACE_CASES)
def test_emitted_runge_kutta_steps_match_reference(emit_tool, disp, eq, solver, tag):
    """Skeleton RK stage/step loops + emitted body, one step at a time from the reference's own
    pre-step states (1e-12) and 5 fused steps (1e-11)."""
    golden_tag = tag
    g = golden("ref_trace_%s_%s_%s" % (disp, tag, solver))
    loose = ILL_CONDITIONED.get((disp, tag), {})
    rec = g["per_step"]
    n = rec.shape[2]
    tag = "rk_%s_%s_%s" % (disp, eq, solver)
    cu, tab, info = emit(emit_tool, disp, eq, solver, tag, dt=float(g["dt"]))
    for step in range(rec.shape[0] - 1):
        out = run_harness(cu, tab, "solver_kernel", rec[step][:8], n, 1, 8, 1, tag)
        for i in range(9):
            ref = rec[step + 1][i]
            if i == 8:
                d_got, d_ref = np.sqrt(out[i]), np.sqrt(ref)
                assert np.max(np.abs(d_got - d_ref)) <= 1.0e-12*np.max(d_ref) + RESIDUAL_FLOOR.get((disp, golden_tag), 1.0e-14)
            else:
                assert rel_dev(out[i], ref) < loose.get(i, 1.0e-12), (step, i, rel_dev(out[i], ref))
    out = run_harness(cu, tab, "solver_kernel", rec[0][:8], n, rec.shape[0] - 1, 8, 1, tag)
    for i in range(8):
        assert rel_dev(out[i], rec[-1][i]) < 10.0*loose.get(i, 1.0e-12), (i, rel_dev(out[i], rec[-1][i]))


def test_emitted_vmec_cold_plasma_dkdt_matches_reference_finite_differences(emit_tool):
    """cold_plasma + VMEC: the reference's symbolic dk/dt is defective on all three components (no closed
    form known), so dk/dt is pinned to the reference's OWN D: 4th-order central differences of D from its
    kernels (tests/golden/ref_fd_cold_plasma_vmec.npz, oracle/make_golden.py vmec_fd).  Stated 1e-7 (the
    accuracy of the differences; observed 7e-9); the reference's symbolic values are 5-16 % away from them."""
    g = golden("ref_fd_cold_plasma_vmec")
    n = g["state"].shape[1]
    cu, tab, info = emit(emit_tool, "cold_plasma", "vmec", "rhs", "rhs_cold_plasma_vmec")
    out = run_harness(cu, tab, "rhs_kernel", g["state"], n, 1, 8, 7, "rhs_cold_plasma_vmec")[8:]
    assert rel_dev(out[6], g["D"]) < 1.0e-12
    for i, fd in enumerate(("dDdx", "dDdy", "dDdz")):
        ref = g[fd]/g["dDdw"]
        assert rel_dev(out[3 + i], ref) < 1.0e-7, (fd, rel_dev(out[3 + i], ref))
        assert rel_dev(g["rhs"][3 + i], ref) > 1.0e-3          # the reference's own symbolic value is not its derivative


def test_emitted_vmec_steps_match_reference(emit_tool):
    """VMEC O-mode (86 Fourier modes, device mode loop with angle recurrence) against a trajectory of the
    reference itself (tests/golden/ref_trace_ordinary_wave_vmec_rk4.npz, oracle/make_golden.py vmec_trace:
    Newton + 20 single steps): Newton root, three single steps from the reference's pre-step states (1e-12),
    20 fused steps (stated 1e-9, observed 6e-16)."""
    g = golden("ref_trace_ordinary_wave_vmec_rk4")
    rec = g["per_step"]
    n = rec.shape[2]
    cu, tab, info = emit(emit_tool, "ordinary_wave", "vmec", "newton", "newton_ordinary_wave_vmec")
    out = run_harness(cu, tab, "loss_kernel", g["state"], n, 1000, 8, 1, "newton_ordinary_wave_vmec", scalar=1.0e-30)
    assert rel_dev(out[5], rec[0][5]) < 1.0e-12
    cu, tab, info = emit(emit_tool, "ordinary_wave", "vmec", "rk4", "rk_ordinary_wave_vmec", dt=float(g["dt"]))
    for step in range(3):
        out = run_harness(cu, tab, "solver_kernel", rec[step][:8], n, 1, 8, 1, "rk_ordinary_wave_vmec")
        for i in range(8):
            assert rel_dev(out[i], rec[step + 1][i]) < 1.0e-12, (step, i, rel_dev(out[i], rec[step + 1][i]))
        assert np.max(np.abs(np.sqrt(out[8]) - np.sqrt(rec[step + 1][8]))) < 1.0e-13
    out = run_harness(cu, tab, "solver_kernel", rec[0][:8], n, 20, 8, 1, "rk_ordinary_wave_vmec")
    for i in range(8):
        assert rel_dev(out[i], rec[20][i]) < 1.0e-9, (i, rel_dev(out[i], rec[20][i]))


@pytest.mark.parametrize("disp,eq,tag", [("extra_ordinary_wave", "efit", "efit"), ("cold_plasma", "slab_density", "slab_density"),
                                         ("cold_plasma", "efit", "efit_interior"), ("cold_plasma", "efit", "efit")])
def test_emitted_newton_matches_reference(emit_tool, disp, eq, tag):
    """Device-resident per-ray Newton skeleton vs the reference's converged kx."""
    g = golden("ref_trace_%s_%s_rk4" % (disp, tag))
    n = g["state"].shape[1]
    name = "newton_%s_%s" % (disp, eq)
    cu, tab, info = emit(emit_tool, disp, eq, "newton", name)
    out = run_harness(cu, tab, "loss_kernel", g["state"], n, 1000, 8, 1, name, scalar=1.0e-30)
    print(disp, tag, "kx deviation", rel_dev(out[5], g["per_step"][0][5]), "max D^2", np.max(out[8]))
    # cold plasma + EFIT: the root is only defined to |noise of D|/|dD/dkx| (a double root in vacuum)
    tol = {("cold_plasma", "efit_interior"): 1.0e-10, ("cold_plasma", "efit"): 1.0e-7}.get((disp, tag), 1.0e-12)
    assert rel_dev(out[5], g["per_step"][0][5]) < tol
    assert np.max(out[8]) < 1.0e-20          # D^2 at the last evaluated iterate


def test_efit_kernel_shape(emit_tool):
    """Structure the design relies on: no divides left, tables grouped by cell, 1-D group staged."""
    cu, tab, info = emit(emit_tool, "cold_plasma", "efit", "rk4", "shape")
    assert info["divides"] == 0 and 0 < info["reciprocals"] <= 16
    assert info["statements"] < 1100          # reference: 3861 statements for the 4-stage step
    cells = sorted(g["cells"] for g in info["groups"])
    assert cells[-1] == 4096 and set(cells[:-1]) == {138}      # the psi(R, Z) spline and the profile splines in psi
    g2d = [g for g in info["groups"] if g["cells"] == 4096][0]
    assert g2d["stride"] == 16 and not g2d["staged"]           # one 128-byte row per cell, read from L1/L2
    for g1d in [g for g in info["groups"] if g["cells"] == 138]:
        assert g1d["staged"] and g1d["stride"] == 4            # staged into shared memory by TMA
    text = open(cu).read()
    # indices: (R, Z) once for the 2-D spline, psi once per profile; clamped on the integer side
    assert text.count("__double2uint_rz(") == 2 + len(cells) - 1 and "fmin(fmax(" not in text
    # psi and its five derivatives come out of ONE pass over the cell's row: 8 16-byte loads, no more
    assert text.count("__ldg(sr") == 8


@pytest.mark.parametrize("disp,eq,kind", [("cold_plasma", "efit", "rk4"), ("extra_ordinary_wave", "efit", "rk4"),
                                          ("extra_ordinary_wave", "efit", "newton"), ("cold_plasma", "slab", "rk2")])
def test_nvrtc_compiles_for_sm_100a(lib, emit_tool, disp, eq, kind):
    """NVRTC needs no GPU: skeleton + emitted body must compile to an sm_100a cubin."""
    cu, tab, info = emit(emit_tool, disp, eq, kind, "nvrtc_%s_%s_%s" % (disp, eq, kind))
    src = open(cu).read().encode()
    cubin, size, log = ctypes.c_void_p(), ctypes.c_size_t(), ctypes.c_void_p()
    rc = lib.gfb_compile_to_cubin(src, None, ctypes.byref(cubin), ctypes.byref(size), ctypes.byref(log))
    msg = ctypes.string_at(log).decode() if log.value else ""
    assert rc == 0, msg
    image = ctypes.string_at(cubin, size.value)
    assert image[:4] == b"\x7fELF" and size.value > 4096
    lib.gfb_free(cubin)
    if log.value:
        lib.gfb_free(log)


#  Absorption.  Tolerances: Im k_amp carries exp(-zeta^2) with zeta^2 up to ~740 and
#  zeta = (1 - w_ce/w)/(n_par v_t/c) is itself a cancellation, so its relative precision is
#  ~zeta^2 * 1e-13; the absolute floor 1e-11 is 1e-14 of |k|.  Re k_amp is only compared where the
#  reference's own value is defined (|zeta| < 26.6: above that erfi overflows, special_functions.hpp:1508).
def absorption_zeta(records):
    import sys
    sys.path.insert(0, ROOT)
    from oracle import port
    from graph_framework_b200.tools.gfbt import read_gfbt
    eq = port.make_equilibrium("efit", read_gfbt(os.path.join(ROOT, "tests", "golden", "efit.gfbt")))
    z = []
    with np.errstate(all="ignore"):
        for r in records:
            s = dict(zip(port.ORDER, r[:8]))
            e = port._expansion_terms(eq, s["w"], (s["kx"], s["ky"], s["kz"]), s["x"], s["y"], s["z"])
            z.append((1.0 - e["ec"]/s["w"])/(e["npara"]*np.sqrt(2.0*port.Q*e["f"]["te"]/port.ME)/port.C))
    return np.array(z)


def assert_kamp_close(re, im, g, zeta, what=""):
    assert np.isfinite(im).all() and np.isfinite(re).all(), what
    assert np.max(np.abs(im - g["kamp_im"]) - 1.0e-8*np.abs(g["kamp_im"])) < 1.0e-11, what
    ok = np.abs(zeta) < 26.6
    assert ok.sum() > 0.4*ok.size
    assert np.max(np.abs(re[ok] - g["kamp_re"][ok])/np.abs(g["kamp_re"][ok])) < 1.0e-10, what


def test_emitted_absorption_kernels_match_reference(emit_tool):
    """weak_damping_kimg_kernel and power (absorption.hpp / xrays.cpp bin_power) on the CPU harness
    against the reference's own JIT kernels (complex<double>, SAFE_MATH) record by record."""
    g = golden("ref_absorb_ordinary_wave_efit")
    rec = g["records"]
    nrec, _, n = rec.shape
    zeta = absorption_zeta(rec)
    cu, tab, info = emit(emit_tool, "none", "efit", "kamp", "kamp_efit")
    assert "gfb::erfi(" in open(cu).read()
    re, im = np.zeros((nrec, n)), np.zeros((nrec, n))
    for j in range(nrec):
        t, w, x, y, z, kx, ky, kz = rec[j][:8]
        out = run_harness(cu, tab, "weak_damping_kimg_kernel", [np.zeros(n), np.zeros(n), kx, ky, kz, x, y, z, t, w],
                          n, 1, 10, 0, "kamp_efit")
        re[j], im[j] = out[0], out[1]
    assert_kamp_close(re, im, g, zeta)
    assert np.max(im) > 5.0                       # the case does cross the resonance

    cu, tab, info = emit(emit_tool, "none", "efit", "power", "power_efit")
    last, power, k_sum = rec[0][2:5].copy(), np.ones(n), np.zeros(n)
    for j in range(1, nrec):
        x, y, z = rec[j][2:5]
        out = run_harness(cu, tab, "power", [x, y, z, last[0], last[1], last[2], g["kamp_im"][j], power, k_sum],
                          n, 1, 9, 1, "power_efit")
        last, power, k_sum = out[3:6].copy(), out[7], out[8]
        assert np.array_equal(last, rec[j][2:5])
        assert np.max(np.abs(power - g["power"][j])) < 1.0e-13
        assert np.max(np.abs(out[9] - g["d_power"][j])) < 1.0e-13
    assert 0.05 < power.min() < 0.6               # most of the power is absorbed by the last record


@pytest.mark.parametrize("kind", ["kamp", "power"])
def test_nvrtc_compiles_absorption_kernels(lib, emit_tool, kind):
    cu, tab, info = emit(emit_tool, "none", "efit", kind, "nvrtc_" + kind)
    cubin, size, log = ctypes.c_void_p(), ctypes.c_size_t(), ctypes.c_void_p()
    rc = lib.gfb_compile_to_cubin(open(cu).read().encode(), None, ctypes.byref(cubin), ctypes.byref(size), ctypes.byref(log))
    assert rc == 0, ctypes.string_at(log).decode() if log.value else ""
    assert ctypes.string_at(cubin, 4) == b"\x7fELF"
    lib.gfb_free(cubin)
    if log.value:
        lib.gfb_free(log)


def test_emitted_index_kernels_match_numpy(lib, emit_tool):
    """index_1D gathers (piecewise.hpp:1436-1640) in the two kernels of graph_pic/xpic.cpp at test
    size, on the CPU harness against a numpy restatement; the indexed arrays are never loaded per ray."""
    n = 64
    rng = np.random.default_rng(4)
    scale, offset, dt = 2.0/63.0, -1.0, 1.0e-2
    x, v, e = rng.normal(0.0, 0.4, n), rng.normal(0.0, 0.25, n), rng.normal(size=n)

    def gather(arr, arg, s, o):
        return arr[np.clip((arg - o)/s, 0, arr.size - 1).astype(int)]

    cu, tab, info = emit(emit_tool, "none", "slab", "pic_push", "pic_push")
    text = open(cu).read()
    assert "v[2] =" not in text and "tg[0][" in text and "group_slot" in text      # e is only indexed
    out = run_harness(cu, tab, "Particle_Push", [x, v, e], n, 1, 3, 0, "pic_push")
    x1, v1 = dt*v, -gather(e, x, scale, offset)
    x2, v2 = dt*(v + v1/2), -gather(e, x + x1/2, scale, offset)
    x3, v3 = dt*(v + v2/2), -gather(e, x + x2/2, scale, offset)
    x4, v4 = dt*(v + v3), -gather(e, x + x3, scale, offset)
    assert np.allclose(out[0], x + (x1 + 2*(x2 + x3) + x4)/6, rtol=1e-14, atol=1e-16)
    assert np.allclose(out[1], v + (v1 + 2*(v2 + v3) + v4)/6, rtol=1e-14, atol=1e-16)
    assert np.array_equal(out[2], e)

    cu, tab, info = emit(emit_tool, "none", "slab", "pic_field", "pic_field")
    grid = scale*np.arange(n) + offset
    field, dens, idx = np.zeros(n), np.zeros(n), np.zeros(n)
    out = run_harness(cu, tab, "Compute_efield", [field, dens, grid, idx, x], n, 3, 5, 0, "pic_field")   # 3 fused steps of 4
    for p in x[:12]:
        d = p - grid
        dens += np.exp(d*d/-0.01)
        field += -1.0/np.exp(d*d/-0.01)*(np.exp(d*d/-0.01)*2.0*d/-0.01)
    assert np.allclose(out[1], dens, rtol=1e-13) and np.allclose(out[0], field, rtol=1e-12, atol=1e-12)
    assert np.array_equal(out[3], np.full(n, 12.0)) and np.array_equal(out[4], x)

    for name in ("pic_push", "pic_field"):
        src = open(os.path.join(BUILD, "emit_%s.cu" % name)).read().encode()
        cubin, size, log = ctypes.c_void_p(), ctypes.c_size_t(), ctypes.c_void_p()
        assert lib.gfb_compile_to_cubin(src, None, ctypes.byref(cubin), ctypes.byref(size), ctypes.byref(log)) == 0
        lib.gfb_free(cubin)


def test_emitted_boris_push_matches_reference(emit_tool):
    """The xkorc step graph (csrc/graph/boris.hpp = graph_korc/xkorc.cpp:66-121) as a fused generic item on the CPU
    harness: 50 pushes in the EFIT field from the reference's start, against the reference's own run
    (tests/golden/ref_korc_efit.npz); the on-axis field is taken from the golden (the device search for it is a GPU test)."""
    g = golden("ref_korc_efit")
    x, y, z, ux, uy, uz = g["start"]
    n = x.size
    gamma = 1.0/np.sqrt(1.0 - (ux*ux + uy*uy + uz*uz))         # the initialize_gamma pre-item
    env_b0 = repr(float(g["b0"]))
    os.environ["GFB_B0"] = env_b0
    try:
        cu, tab, info = emit(emit_tool, "none", "efit", "boris", "boris_efit", dt=0.5)
    finally:
        os.environ.pop("GFB_B0", None)
    out = run_harness(cu, tab, "step", [x, y, z, gamma*ux, gamma*uy, gamma*uz, gamma], n, 50, 7, 0, "boris_efit")
    for i, name in enumerate(("x", "y", "z", "ux", "uy", "uz", "gamma")):
        assert rel_dev(out[i], g["end"][i]) < 1.0e-10, (name, rel_dev(out[i], g["end"][i]))
    assert info["divides"] == 0 and info["reciprocals"] <= 3


def test_horner_chain_construction_still_steps_like_the_reference(emit_tool):
    """GFB_SPLINE_NODES=0: EFIT built from Horner chains of piecewise nodes (the reference's build_1D_spline /
    build_psi shape) instead of the spline node families -- the A/B switch of DESIGN.md section 2 must keep working:
    same per-step parity, and the family version of the same kernel has fewer statements."""
    g = golden("ref_trace_extra_ordinary_wave_efit_rk4")
    rec = g["per_step"]
    n = rec.shape[2]
    os.environ["GFB_SPLINE_NODES"] = "0"
    try:
        cu, tab, chains = emit(emit_tool, "extra_ordinary_wave", "efit", "rk4", "rk_xmode_chains", dt=float(g["dt"]))
    finally:
        os.environ.pop("GFB_SPLINE_NODES", None)
    assert "fma(v[4], c" in open(cu).read() or "p0_0" in open(cu).read()         # piecewise groups, not spline rows
    out = run_harness(cu, tab, "solver_kernel", rec[0][:8], n, 1, 8, 1, "rk_xmode_chains")
    for i in range(8):
        assert rel_dev(out[i], rec[1][i]) < 1.0e-12, (i, rel_dev(out[i], rec[1][i]))
    _, _, families = emit(emit_tool, "extra_ordinary_wave", "efit", "rk4", "rk_xmode_families", dt=float(g["dt"]))
    assert families["statements"] < 0.8*chains["statements"]
