#!/usr/bin/env python3
"""Generate golden input/output vectors from the reference itself (oracle/_ref/ref_driver).

TEST INFRASTRUCTURE.  Run in the build container (needs /root/reference to have been compiled
into oracle/_ref by `make -C oracle ref`).  The .npz files written to tests/golden/ are
committed so the CPU and GPU test-suites never need /root/reference at run time.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)
from oracle import reference                    # noqa: E402
from graph_framework_b200 import workloads      # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")


def save(name, **arrays):
    path = os.path.join(OUT, name + ".npz")
    np.savez_compressed(path, **arrays)
    print("wrote", path, os.path.getsize(path), "bytes")


def rhs_cases():
    n = 64
    cases = [("ordinary_wave", "efit"), ("extra_ordinary_wave", "efit"), ("cold_plasma", "efit"),
             ("cold_plasma", "slab"), ("cold_plasma", "slab_density"), ("ordinary_wave", "slab_density"),
             ("bohm_gross", "no_magnetic_field"), ("simple", "slab"), ("cold_plasma", "gaussian_density")]
    for disp, eq in cases:
        if eq == "efit":
            s = workloads.interior_states(n, seed=1)
        else:
            s = workloads.slab_ensemble(n, seed=2)
            if eq == "gaussian_density":
                s["x"] = s["x"]*0.5
        out = reference.rhs(disp, eq, s)
        save("ref_rhs_%s_%s" % (disp, eq), state=workloads.pack(s), rhs=out)


def trace_cases():
    n = 32
    cases = [("extra_ordinary_wave", "efit", 2.0e-5, "kx", "rk4"),
             ("ordinary_wave", "efit", 2.0e-5, "kx", "rk4"),
             ("cold_plasma", "efit", 2.0e-5, "kx", "rk4"),
             ("cold_plasma", "slab_density", 1.0e-3, "kx", "rk4"),
             ("ordinary_wave", "slab_density", 1.0e-3, "kx", "rk4"),
             ("cold_plasma", "slab", 1.0e-3, "kx", "rk2"),
             ("simple", "slab", 1.0e-2, "kx", "rk4")]
    for disp, eq, dt, init, solver in cases:
        s = workloads.efit_ensemble(n, seed=5) if eq == "efit" else workloads.slab_ensemble(n, seed=6)
        if disp == "simple":
            s["w"][:] = 1.0
            s["kx"][:] = 0.6
            s["ky"] = np.abs(s["ky"])*0.01 + 0.2
            s["kz"] = np.abs(s["kz"])*0.01 + 0.2
        per_step = reference.trace(disp, eq, s, dt, 5, save_every=1, init=init, solver=solver)
        long = reference.trace(disp, eq, s, dt, 200, save_every=100, init=init, solver=solver)
        save("ref_trace_%s_%s_%s" % (disp, eq, solver), state=workloads.pack(s), dt=np.array(dt),
             per_step=per_step, long=long)


def interior_case():
    """cold_plasma + EFIT started INSIDE the plasma (the efit_example rays of trace_cases start at
    R = 2.5 in vacuum, where cold-plasma D is doubly degenerate and the reference's dkz/dt is a
    ratio of two rounding-level numbers): Newton for kx, 5 single steps, 200 steps in two blocks."""
    n, dt = 32, 2.0e-5
    s = workloads.interior_states(n, seed=9)
    per_step = reference.trace("cold_plasma", "efit", s, dt, 5, save_every=1, init="kx", solver="rk4")
    long = reference.trace("cold_plasma", "efit", s, dt, 200, save_every=100, init="kx", solver="rk4")
    save("ref_trace_cold_plasma_efit_interior_rk4", state=workloads.pack(s), dt=np.array(dt), per_step=per_step, long=long)


def bench_case():
    # xrays_bench initial conditions (every ray identical), xrays_bench.cpp:62-79
    s = workloads.bench_rays(8)
    rec = reference.trace("cold_plasma", "efit", s, 1.0e-3, 10, save_every=5, init="kx")
    save("ref_trace_xrays_bench", state=workloads.pack(s), dt=np.array(1.0e-3), records=rec)


def korc_case():
    n = 16
    x, y, z, ux, uy, uz = workloads.boris_ensemble(n, seed=4)
    out, info = reference.korc("efit", x, y, z, ux, uy, uz, 50)
    save("ref_korc_efit", start=np.stack([x, y, z, ux, uy, uz]), end=out,
         b0=np.array(info["b0"]), larmor_radius=np.array(info["larmor_radius"]))


def vmec_case():
    """VMEC right-hand sides from the reference (graph build ~10 min, kernel compile minutes)."""
    n = 16
    s = workloads.vmec_states(n, seed=8)
    for disp in ("ordinary_wave", "cold_plasma"):
        out = reference.rhs(disp, "vmec", s)
        save("ref_rhs_%s_vmec" % disp, state=workloads.pack(s), rhs=out)


def vmec_trace_case(disp="ordinary_wave"):
    """VMEC trajectory from the reference: Newton solve for k_s, then 20 single RK4 steps, every
    state recorded (records 0..3 feed the per-step tests, record 20 the block test).  One run costs
    the reference ~10 min of graph building plus the g++ compile of its 54 k-statement kernel."""
    n, dt = 16, 1.0e-4
    s = workloads.vmec_states(n, seed=8)
    rec = reference.trace(disp, "vmec", s, dt, 20, save_every=1, init="kx", solver="rk4")
    save("ref_trace_%s_vmec_rk4" % disp, state=workloads.pack(s), dt=np.array(dt), per_step=rec)


def vmec_fd_case():
    """cold_plasma + VMEC: the reference's symbolic dk/dt is defective on all three components (the
    same reducer rule as for EFIT, no closed form known here), so this back end's dk/dt is pinned to the
    reference's OWN D instead: D at the base states and at +-h, +-2h along s, u, v and w (4th-order central
    differences), all in ONE reference run (the VMEC graph build dominates its cost)."""
    n, h = 16, 1.0e-5
    s = workloads.vmec_states(n, seed=8)
    batch = {k: [np.asarray(s[k], dtype=np.float64)] for k in workloads.ORDER}
    order = []
    for var in ("x", "y", "z", "w"):
        for mult in (-2.0, -1.0, 1.0, 2.0):
            order.append((var, mult))
            for k in workloads.ORDER:
                batch[k].append(s[k] + (mult*h if k == var else 0.0))
    stacked = {k: np.concatenate(v) for k, v in batch.items()}
    out = reference.rhs("cold_plasma", "vmec", stacked)
    D = out[6].reshape(len(order) + 1, n)
    fd = {}
    for j, var in enumerate(("x", "y", "z", "w")):
        m2, m1, p1, p2 = (D[1 + 4*j + i] for i in range(4))
        fd[var] = (m2 - 8.0*m1 + 8.0*p1 - p2)/(12.0*h)
    save("ref_fd_cold_plasma_vmec", state=workloads.pack(s), h=np.array(h), D=D[0], rhs=out[:, :n],
         dDdx=fd["x"], dDdy=fd["y"], dDdz=fd["z"], dDdw=fd["w"])


def f3_cases():
    """split_simplextic (solver.hpp:1017-1130) on the two separable Hamiltonians, and the first
    steps of adaptive_rk4 (solver.hpp:882-1006).  The reference's adaptive rule -- Newton on (dt, lambda) of
    1/dt + lambda D_next^2 -- is ill-posed: its own runs give |dt| ~ 1e13 or NaN from the second step on, so
    only the first step (where some rays are finite) is recorded."""
    n = 32
    s = workloads.slab_ensemble(n, seed=6)
    for disp in ("bohm_gross", "light_wave"):
        per_step = reference.trace(disp, "no_magnetic_field", s, 1.0e-3, 5, save_every=1, init="kx", solver="split_simplextic")
        long = reference.trace(disp, "no_magnetic_field", s, 1.0e-3, 200, save_every=100, init="kx", solver="split_simplextic")
        save("ref_trace_%s_no_magnetic_field_split_simplextic" % disp, state=workloads.pack(s), dt=np.array(1.0e-3),
             per_step=per_step, long=long)
    a = workloads.slab_ensemble(n, seed=6)
    a["w"][:] = 900.0
    a["kx"][:] = 1000.0
    a["ky"][:] = 0.25
    a["kz"][:] = 0.15
    a["x"] = np.linspace(-0.2, 0.2, n)
    a["y"][:] = 0.0
    a["z"][:] = 0.0
    rec = reference.trace_adaptive("cold_plasma", "gaussian_density", a, 0.5e-4, 2, save_every=1, init="kx")
    save("ref_trace_cold_plasma_gaussian_density_adaptive_rk4", state=workloads.pack(a), dt=np.array(0.5e-4), records=rec)


def cells_case():
    """Which cell the reference's compiled kernels select within a few ulp of every cell edge of the EFIT
    R and Z grids (index work must be bit-exact: one cell off is a different polynomial)."""
    from graph_framework_b200.tools.gfbt import read_gfbt
    t = read_gfbt(os.path.join(OUT, "efit.gfbt"))
    out = {}
    for tag, off, scale, n in (("r", float(np.ravel(t["rmin"])[0]), float(np.ravel(t["dr"])[0]), 64),
                               ("z", float(np.ravel(t["zmin"])[0]), float(np.ravel(t["dz"])[0]), 64),
                               ("psi", float(np.ravel(t["psimin"])[0]), float(np.ravel(t["dpsi"])[0]), 138)):
        edges = off + scale*np.arange(-1, n + 2)
        xs = [edges]
        up, down = edges.copy(), edges.copy()
        for _ in range(3):
            up, down = np.nextafter(up, np.inf), np.nextafter(down, -np.inf)
            xs += [up.copy(), down.copy()]
        rng = np.random.default_rng(17)
        xs.append(rng.uniform(off - scale, off + (n + 1)*scale, 2000))
        x = np.concatenate(xs)
        out["x_" + tag] = x
        out["grid_" + tag] = np.array([scale, off, n])
        out["cell_" + tag] = reference.cells(x, scale, off, n)
    save("ref_cells_efit", **out)


def defect_case():
    """Evidence for the reference's symbolic dD/dz defect (cold_plasma in a z-dependent field):
    its own D at z +- h and w +- h next to its own symbolic dkz/dt."""
    n = 16
    s = workloads.interior_states(n, seed=7)
    h = 1.0e-6
    base = reference.rhs("cold_plasma", "efit", s)
    out = {"state": workloads.pack(s), "rhs": base, "h": np.array(h)}
    for var in ("z", "w", "x"):
        for sign, tag in ((1.0, "p"), (-1.0, "m")):
            p = dict(s)
            p[var] = s[var] + sign*h
            out["D_%s_%s" % (var, tag)] = reference.rhs("cold_plasma", "efit", p)[6]
    save("ref_defect_cold_plasma_efit", **out)


def absorb_case():
    """The absorption stages of xrays on an O-mode trajectory that crosses the electron cyclotron
    fundamental of the EFIT case (w = 700: resonance near R = 2.06), and erfi on a grid."""
    n = 32
    s = workloads.efit_ensemble(n, seed=0)
    records = reference.trace("ordinary_wave", "efit", s, 1.0e-3, 600, save_every=20, init="kx", solver="rk4")
    out = reference.absorb("efit", records[:, :8])
    save("ref_absorb_ordinary_wave_efit", state=workloads.pack(s), dt=np.array(1.0e-3), sub_steps=np.array(20),
         records=records, **out)
    rng = np.random.default_rng(11)
    x = np.concatenate([[0.0, 1e-300, 1e-8, 0.01, 0.0308, 0.0309, 0.031, 0.5, 1.0, 26.6, 26.8, 26.9, 44.9, 45.0, 45.1,
                         100.0, 1e6, 6e7, -1.3, -30.0], rng.uniform(-46, 46, 600), rng.uniform(-3, 3, 400)])
    w_im, erfi = reference.erfi(x)
    save("ref_erfi", x=x, w_im=w_im, erfi=erfi)


if __name__ == "__main__":
    if not reference.available():
        raise SystemExit("oracle/_ref/ref_driver missing: run `make -C oracle ref` first")
    which = sys.argv[1:] or ["rhs", "trace", "bench", "korc", "defect", "interior"]
    if "rhs" in which:
        rhs_cases()
    if "trace" in which:
        trace_cases()
    if "bench" in which:
        bench_case()
    if "f3" in which:
        f3_cases()
    if "cells" in which:
        cells_case()
    if "interior" in which:
        interior_case()
    if "korc" in which:
        korc_case()
    if "defect" in which:
        defect_case()
    if "vmec" in which:
        vmec_case()
    if "absorb" in which:
        absorb_case()
    if "vmec_fd" in which:
        vmec_fd_case()
    if "vmec_trace" in which:
        vmec_trace_case("ordinary_wave")
    if "vmec_trace_cold" in which:
        vmec_trace_case("cold_plasma")
