"""Python front for the reference oracle binary.  TEST INFRASTRUCTURE.

oracle/_ref/ref_driver is the UNMODIFIED reference graph + solver code compiled
with the g++ stand-in context (see oracle/Makefile, oracle/ref_driver.cpp).  Only
tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
legs may import this module; the product never does.
"""
import json
import os
import subprocess
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
DRIVER = os.path.join(HERE, "_ref", "ref_driver")
ORDER = ("t", "w", "x", "y", "z", "kx", "ky", "kz")


def available():
    return os.path.exists(DRIVER) and os.access(DRIVER, os.X_OK)


def _run(args, timeout=7200, threads=None):
    env = dict(os.environ)
    env.setdefault("GFB_EFIT_FILE", os.path.join(ROOT, "tests", "golden", "efit.gfbt"))
    env.setdefault("GFB_VMEC_FILE", os.path.join(ROOT, "tests", "golden", "vmec.gfbt"))
    if threads:
        env["GFB_ORACLE_THREADS"] = str(threads)
#  Complex kernels (absorb) include the reference's special_functions.hpp: only where it exists.
    if os.path.isdir("/root/reference/graph_framework"):
        env.setdefault("GFB_REFERENCE_INCLUDE", "/root/reference/graph_framework")
    return subprocess.run([DRIVER] + [str(a) for a in args], cwd=ROOT, env=env, timeout=timeout,
                          check=True, capture_output=True, text=True).stdout


def _pack(state):
    return np.stack([np.broadcast_to(np.asarray(state[k], dtype=np.float64), np.shape(state["w"]))
                     for k in ORDER])


def trace(dispersion, equilibrium, state, dt, nsteps, save_every=None, init="kx", solver="rk4"):
    """Returns an array [records, 9, N]: record 0 = state after init, then every `save_every` steps.
    Rows: t, w, x, y, z, kx, ky, kz, residual."""
    arr = _pack(state)
    n = arr.shape[1]
    save_every = save_every or nsteps
    with tempfile.TemporaryDirectory() as d:
        fin, fout = os.path.join(d, "in.bin"), os.path.join(d, "out.bin")
        arr.tofile(fin)
        _run(["trace", dispersion, equilibrium, solver, n, repr(float(dt)), nsteps, save_every, init or "none", fin, fout])
        return np.fromfile(fout).reshape(-1, 9, n)


def trace_adaptive(dispersion, equilibrium, state, dt, nsteps, save_every=1, init="kx"):
    """solver::adaptive_rk4: records [records, 10, N] = t, w, x, y, z, kx, ky, kz, residual, dt (per ray)."""
    arr = _pack(state)
    n = arr.shape[1]
    with tempfile.TemporaryDirectory() as d:
        fin, fout = os.path.join(d, "in.bin"), os.path.join(d, "out.bin")
        arr.tofile(fin)
        _run(["trace", dispersion, equilibrium, "adaptive_rk4", n, repr(float(dt)), nsteps, save_every, init or "none", fin, fout])
        return np.fromfile(fout).reshape(-1, 10, n)


def rhs(dispersion, equilibrium, state):
    """dxdt, dydt, dzdt, dkxdt, dkydt, dkzdt, D  as a [7, N] array."""
    arr = _pack(state)
    n = arr.shape[1]
    with tempfile.TemporaryDirectory() as d:
        fin, fout = os.path.join(d, "in.bin"), os.path.join(d, "out.bin")
        arr.tofile(fin)
        _run(["rhs", dispersion, equilibrium, n, fin, fout])
        return np.fromfile(fout).reshape(7, n)


def bench(dispersion, equilibrium, n, dt, nsteps, threads, state=None, blocks=1):
    """Times the reference stepping loop (xrays_bench.cpp:88-102) on `threads` host threads;
    `blocks` repetitions of nsteps after one setup are timed separately (result["block_s"])."""
    with tempfile.TemporaryDirectory() as d:
        fin = "-"
        if state is not None:
            fin = os.path.join(d, "in.bin")
            _pack(state).tofile(fin)
        out = _run(["bench", dispersion, equilibrium, n, repr(float(dt)), nsteps, threads, fin, blocks])
    return json.loads(out.strip().splitlines()[-1])


def korc(equilibrium, x, y, z, ux, uy, uz, nsteps):
    arr = np.stack([np.asarray(v, dtype=np.float64) for v in (x, y, z, ux, uy, uz)])
    n = arr.shape[1]
    with tempfile.TemporaryDirectory() as d:
        fin, fout = os.path.join(d, "in.bin"), os.path.join(d, "out.bin")
        arr.tofile(fin)
        out = _run(["korc", equilibrium, n, nsteps, fin, fout])
        return np.fromfile(fout).reshape(7, n), json.loads(out.strip().splitlines()[-1])


def absorb(equilibrium, records):
    """records [nrec, 8, n] in the order t,w,x,y,z,kx,ky,kz -> dict of [nrec, n] arrays
    kamp_re, kamp_im, power, d_power (ref_driver absorb: weak_damping then bin_power)."""
    arr = np.ascontiguousarray(records, dtype=np.float64)
    nrec, _, n = arr.shape
    with tempfile.TemporaryDirectory() as d:
        fin, fout = os.path.join(d, "in.bin"), os.path.join(d, "out.bin")
        arr.tofile(fin)
        _run(["absorb", equilibrium, n, nrec, fin, fout])
        out = np.fromfile(fout).reshape(nrec, 4, n)
    return {"kamp_re": out[:, 0], "kamp_im": out[:, 1], "power": out[:, 2], "d_power": out[:, 3]}


def erfi(x):
    """special::w_im(x), Re special::erfi(x + 0i) of the reference's special_functions.hpp."""
    arr = np.ascontiguousarray(x, dtype=np.float64)
    with tempfile.TemporaryDirectory() as d:
        fin, fout = os.path.join(d, "in.bin"), os.path.join(d, "out.bin")
        arr.tofile(fin)
        _run(["erfi", arr.size, fin, fout])
        out = np.fromfile(fout).reshape(2, arr.size)
    return out[0], out[1]


def reducer_defect():
    """The four-variable reproduction of the reducer rule behind the reference's wrong cold-plasma
    dD/dz (ref_driver reducer); tests/golden/ref_reducer_defect.json is its committed output."""
    return json.loads(_run(["reducer"]).strip().splitlines()[-1])


def cells(x, scale, offset, ncells):
    """Table cell the reference's compiled kernel selects for each argument (ref_driver cells)."""
    arr = np.ascontiguousarray(x, dtype=np.float64)
    with tempfile.TemporaryDirectory() as d:
        fin, fout = os.path.join(d, "in.bin"), os.path.join(d, "out.bin")
        arr.tofile(fin)
        _run(["cells", arr.size, repr(float(scale)), repr(float(offset)), ncells, fin, fout])
        return np.fromfile(fout)
