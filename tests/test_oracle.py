"""Pin the oracle before trusting it (no GPU needed).

1. The numpy restatement (oracle/port.py) against the reference's only golden file,
   graph_tests/efit_gold.nc, with the reference's own tolerances (efit_test.cpp:131-187).
2. The restatement against outputs of the reference itself (tests/golden/ref_*.npz written by
   oracle/make_golden.py from oracle/_ref/ref_driver): right-hand sides, single steps, Newton roots.
3. Evidence for the one place where the reference is NOT used as the expected value: its symbolic
   dD/dz for cold_plasma in a z-dependent field.
"""
import os

import numpy as np
import pytest

from conftest import GOLDEN, golden, rel_dev, rel_devs, assert_rhs_close

ORDER = ("t", "w", "x", "y", "z", "kx", "ky", "kz")
RHS = ("dxdt", "dydt", "dzdt", "dkxdt", "dkydt", "dkzdt", "D")


def unpack(a):
    return {k: np.array(a[i]) for i, k in enumerate(ORDER)}


@pytest.fixture(scope="module")
def efit(efit_tables):
    from oracle import port
    return port.Efit(efit_tables)


def test_fixture_tables_match_reference_files(efit_tables):
    """When the reference tree is present (build container) the committed GFBT fixture must be
    byte-equal to a fresh conversion of graph_tests/efit.nc."""
    nc = "/root/reference/graph_tests/efit.nc"
    if not os.path.exists(nc):
        pytest.skip("reference tree not present")
    from graph_framework_b200.tools.h5lite import H5File
    h = H5File(nc)
    for name in ("psi_c00", "psi_c33", "fpol_c2", "ne_c3", "te_c0", "pressure_c1", "rmin", "dpsi"):
        assert np.array_equal(h.read(name), efit_tables[name].reshape(h.read(name).shape)), name
    assert efit_tables["psi_c00"].shape == (64, 64) and efit_tables["te_c0"].shape == (138,)
    assert np.array_equal(efit_tables["ne_c0"], efit_tables["te_c0"])


def test_port_matches_efit_gold(efit):
    """efit_test.cpp:131-187: B, ne, te on the 51 x 101 (R, Z) grid at y = 0; tolerances on the
    squared relative error 4e-12 (B), 1e-12 (ne), 5e-13... the reference uses per-quantity bounds;
    the loosest that it accepts is asserted here for each."""
    from graph_framework_b200.tools.gfbt import read_gfbt
    g = read_gfbt(os.path.join(GOLDEN, "efit_gold.gfbt"))
    R, Z = np.meshgrid(g["r_grid"], g["z_grid"], indexing="ij")
    with np.errstate(all="ignore"):
        f = efit.fields(R.ravel().astype(complex), np.zeros(R.size, complex), Z.ravel().astype(complex))
    checks = [("bx_grid", f["b"][0], 4.0e-12), ("by_grid", f["b"][1], 4.0e-12), ("bz_grid", f["b"][2], 4.0e-12),
              ("ne_grid", f["ne"], 1.0e-12), ("te_grid", f["te"], 1.0e-12)]
    for name, val, tol in checks:
        gold = g[name].ravel()
        v = val.real
        m = np.abs(gold) > 0
        assert np.max(((v[m] - gold[m])/gold[m])**2) < tol, name
        assert np.max(np.abs(v[~m])) == 0.0 if (~m).any() else True


@pytest.mark.parametrize("disp,eq", [("ordinary_wave", "efit"), ("extra_ordinary_wave", "efit"),
                                     ("cold_plasma", "slab"), ("cold_plasma", "slab_density"),
                                     ("ordinary_wave", "slab_density"), ("bohm_gross", "no_magnetic_field"),
                                     ("simple", "slab"), ("cold_plasma", "gaussian_density")])
def test_port_rhs_matches_reference(efit, disp, eq):
    from oracle import port
    g = golden("ref_rhs_%s_%s" % (disp, eq))
    e = efit if eq == "efit" else port.make_equilibrium(eq)
    with np.errstate(all="ignore"):
        got = port.rhs(disp, e, unpack(g["state"]))
    for i, k in enumerate(RHS):
        if np.max(np.abs(g["rhs"][i])) == 0.0:
            assert np.max(np.abs(got[k])) < 1.0e-300
            continue
        assert_rhs_close(got[k], g["rhs"][i], (disp, eq, k))


@pytest.mark.parametrize("disp,eq,solver", [("extra_ordinary_wave", "efit", "rk4"), ("ordinary_wave", "efit", "rk4"),
                                            ("cold_plasma", "slab_density", "rk4"), ("cold_plasma", "slab", "rk2")])
def test_port_single_steps_match_reference(efit, disp, eq, solver):
    from oracle import port
    g = golden("ref_trace_%s_%s_%s" % (disp, eq, solver))
    rec = g["per_step"]
    e = efit if eq == "efit" else port.make_equilibrium(eq)
    step = port.rk4_step if solver == "rk4" else port.rk2_step
    for s in range(rec.shape[0] - 1):
        with np.errstate(all="ignore"):
            nxt, res = step(disp, e, unpack(rec[s][:8]), float(g["dt"]))
        for i, k in enumerate(ORDER):
            assert rel_dev(nxt[k], rec[s + 1][i]) < 1.0e-12, (s, k)
        assert np.max(np.abs(res - rec[s + 1][8])) <= 1.0e-10*np.max(np.abs(rec[s + 1][8])) + 1.0e-28


@pytest.mark.parametrize("per_ray", [True, False])
def test_port_newton_matches_reference(efit, per_ray):
    from oracle import port
    g = golden("ref_trace_extra_ordinary_wave_efit_rk4")
    with np.errstate(all="ignore"):
        s = port.newton("extra_ordinary_wave", efit, unpack(g["state"]), "kx", per_ray=per_ray)
    assert rel_dev(s["kx"], g["per_step"][0][5]) < 1.0e-12


def test_reference_dkz_defect(efit):
    """The reference's symbolic dkz/dt for cold_plasma + EFIT disagrees with a central finite
    difference of the reference's OWN D(z +- h)/D(w +- h) by orders of magnitude, while its dkx/dt
    agrees; the port agrees with the finite difference.  (Root: the reference's reduction of
    `b_hat->cross(n)->length()`; ordinary/extra_ordinary_wave, which use nperp->dot(nperp), are
    fine.)  This is why dkz/dt of that one case is checked against the port instead."""
    from oracle import port
    g = golden("ref_defect_cold_plasma_efit")
    h = float(g["h"])
    dDdw = (g["D_w_p"] - g["D_w_m"])/(2*h)
    fd_dkz = (g["D_z_p"] - g["D_z_m"])/(2*h)/dDdw
    fd_dkx = (g["D_x_p"] - g["D_x_m"])/(2*h)/dDdw
    ref_dkx, ref_dkz = g["rhs"][3], g["rhs"][5]
    inside = np.abs(fd_dkx) > 1.0e-3*np.max(np.abs(fd_dkx))        # rays in the plasma
    assert inside.sum() >= 8
    assert np.max(np.abs(ref_dkx[inside] - fd_dkx[inside])/np.abs(fd_dkx[inside])) < 1.0e-4
    assert np.median(np.abs(ref_dkz[inside] - fd_dkz[inside])/np.abs(fd_dkz[inside])) > 1.0       # the defect
    with np.errstate(all="ignore"):
        mine = port.rhs("cold_plasma", efit, unpack(g["state"]), reference_defects=False)
    assert np.max(np.abs(mine["dkzdt"][inside] - fd_dkz[inside])/np.maximum(np.abs(fd_dkz[inside]), 1e-3*np.max(np.abs(fd_dkz)))) < 1.0e-3
    assert np.max(rel_devs(mine["dkxdt"], ref_dkx)) < 1.0e-6


def test_reference_reducer_defect_is_pinned(efit):
    """Root cause of the dkz defect, localised: the reference's reducer turns ((A W)^2 B)/(C^2 W^4)
    into A^2 B/C^4 (fixture written by `ref_driver reducer`; re-run here when oracle/_ref exists).
    The port's closed form for what that does to cold_plasma + EFIT (cold_plasma_reference_defect)
    reproduces the reference's own dkz/dt on 80 states."""
    import json
    import os
    from conftest import GOLDEN
    from oracle import port, reference
    with open(os.path.join(GOLDEN, "ref_reducer_defect.json")) as f:
        fix = json.load(f)
    runs = [fix] + ([reference.reducer_defect()] if reference.available() else [])
    for r in runs:
        assert abs(r["reference_graph"] - r["a2b_over_c4"]) < 1.0e-15*abs(r["a2b_over_c4"])       # what it became
        assert abs(r["reference_graph"]/r["direct"]) > 1.0e4                                        # not what it is
        assert abs(r["reference_df"]/r["central_difference"]) > 1.0e4
    for name in ("ref_rhs_cold_plasma_efit", "ref_defect_cold_plasma_efit"):
        g = golden(name)
        with np.errstate(all="ignore"):
            mine = port.rhs("cold_plasma", efit, unpack(g["state"]))
            true = port.rhs("cold_plasma", efit, unpack(g["state"]), reference_defects=False)
        d = rel_devs(mine["dkzdt"], g["rhs"][5])
        assert np.median(d) < 1.0e-13 and np.max(d) < 1.0e-8, (name, np.median(d), np.max(d))
        assert np.median(rel_devs(true["dkzdt"], g["rhs"][5])) > 0.1


def test_port_steps_cold_plasma_like_the_reference(efit):
    """With the defect restated, the port follows the reference's cold-plasma + EFIT trajectory inside the
    plasma step by step (1e-12), which nothing could before."""
    from oracle import port
    g = golden("ref_trace_cold_plasma_efit_interior_rk4")
    rec = g["per_step"]
    with np.errstate(all="ignore"):
        for step in range(rec.shape[0] - 1):
            out, res = port.rk4_step("cold_plasma", efit, unpack(rec[step][:8]), float(g["dt"]))
            for i, k in enumerate(port.ORDER):
                assert rel_dev(out[k], rec[step + 1][i]) < 1.0e-12, (step, k)


def test_port_deposit_matches_numpy_histogram():
    """utilities/bin.py bins with tf.histogram-like half-open uniform bins; cross-check with numpy."""
    from oracle import port
    rng = np.random.default_rng(3)
    n = 5000
    x, y, z, w = rng.uniform(-1.2, 1.2, n), rng.uniform(-1.2, 1.2, n), rng.uniform(-1.2, 1.2, n), rng.uniform(0, 1, n)
    lo, hi, bins = (-1.0, -1.0, -1.0), (1.0, 1.0, 1.0), (8, 9, 10)
    h = port.deposit(x, y, z, w, lo, hi, bins)
    ok = (np.abs(x) < 1) & (np.abs(y) < 1) & (np.abs(z) < 1)
    ref, _ = np.histogramdd(np.stack([x[ok], y[ok], z[ok]], 1), bins=bins, range=list(zip(lo, hi)), weights=w[ok])
    assert np.allclose(h, ref, rtol=1e-12)


def test_port_erfi_matches_reference():
    """special::erfi for real arguments (special_functions.hpp:1504-1512, 545-562): the port's scipy
    based evaluation against values produced by the reference's own header."""
    from oracle import port
    g = golden("ref_erfi")
    mine = port.erfi_real(g["x"])
    fin = np.isfinite(g["erfi"])
    assert np.array_equal(np.isinf(mine), np.isinf(g["erfi"]))
    assert rel_dev(mine[fin], g["erfi"][fin]) < 5.0e-15


def test_port_absorption_matches_reference(efit):
    """Weak damping k_amp and the power stage against the reference's own JIT kernels
    (complex<double>, SAFE_MATH; oracle/_ref absorb mode).  Im k_amp carries exp(-zeta^2): relative
    1e-8 with an absolute floor of 1e-11 (1e-14 of |k|); Re k_amp only where the reference value is
    defined (|zeta| < 26.6, below the overflow of erfi)."""
    from oracle import port
    g = golden("ref_absorb_ordinary_wave_efit")
    rec = g["records"]
    k = np.array([port.weak_damping(efit, unpack(r[:8])) for r in rec])
    assert np.isfinite(k.real).all() and np.isfinite(k.imag).all()
    assert np.max(np.abs(k.imag - g["kamp_im"]) - 1.0e-8*np.abs(g["kamp_im"])) < 1.0e-11
    with np.errstate(all="ignore"):
        zeta = []
        for r in rec:
            s = unpack(r[:8])
            e = port._expansion_terms(efit, s["w"], (s["kx"], s["ky"], s["kz"]), s["x"], s["y"], s["z"])
            zeta.append((1.0 - e["ec"]/s["w"])/(e["npara"]*np.sqrt(2.0*port.Q*e["f"]["te"]/port.ME)/port.C))
    ok = np.abs(np.array(zeta)) < 26.6
    assert ok.mean() > 0.4
    assert rel_dev(k.real[ok], g["kamp_re"][ok]) < 1.0e-11
    power, d_power = port.power_stage(rec[:, 2:5], g["kamp_im"])
    assert np.max(np.abs(power - g["power"])) < 1.0e-13
    assert np.max(np.abs(d_power - g["d_power"])) < 1.0e-13
    assert np.median(g["power"][-1]) < 0.6 and g["kamp_im"].max() > 5.0
