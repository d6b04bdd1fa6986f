import os
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200; run with `-m gpu` on the GPU box")


@pytest.fixture(scope="session")
def lib():
    """libgfb200.so, built on demand (make) -- the library itself, never a substitute."""
    so = os.path.join(ROOT, "graph_framework_b200", "libgfb200.so")
    if not os.path.exists(so):
        subprocess.run(["make", "-j4"], cwd=ROOT, check=True)
        subprocess.run(["make", "-j2", "tests"], cwd=ROOT, check=True)
    from graph_framework_b200 import _lib
    return _lib.lib


@pytest.fixture(scope="session")
def efit_tables():
    from graph_framework_b200.tools.gfbt import read_gfbt
    return read_gfbt(os.path.join(GOLDEN, "efit.gfbt"))


def golden(name):
    return np.load(os.path.join(GOLDEN, name + ".npz"))


def rel_dev(got, ref, floor=1.0e-9):
    """Max relative deviation with an absolute floor of `floor` x the ensemble scale, for
    components that are ~0 on some rays (SURVEY.md 8d parity harness)."""
    got = np.asarray(got, dtype=np.float64)
    ref = np.asarray(ref, dtype=np.float64)
    scale = np.maximum(np.abs(ref), floor*max(float(np.max(np.abs(ref))), 1.0e-300))
    return float(np.max(np.abs(got - ref)/scale))


def rel_devs(got, ref, floor=1.0e-9):
    got = np.asarray(got, dtype=np.float64)
    ref = np.asarray(ref, dtype=np.float64)
    scale = np.maximum(np.abs(ref), floor*max(float(np.max(np.abs(ref))), 1.0e-300))
    return np.abs(got - ref)/scale


def assert_rhs_close(got, ref, what, median_tol=1.0e-12, max_tol=1.0e-6):
    """Right-hand-side components are compared with two bounds.  The reference evaluates its
    splines with coefficients folded into physical coordinates (equilibrium.hpp:1121-1133): psi ~ 0.3
    is a sum of terms up to ~4e7, so ANY re-association (the reference's own kernels are compiled
    with -ffast-math) moves individual rays by up to ~1e-8 relative in dk/dt.  Three independent
    evaluations (reference, numpy port, this back end) differ pairwise by that much on the same few
    rays and by ~1e-14 on the rest, hence: median <= 1e-12, worst ray <= 1e-6.  The per-STEP parity
    tests keep the north star's 1e-12 because the increments are scaled by dt."""
    d = rel_devs(got, ref)
    assert float(np.median(d)) <= median_tol, (what, "median", float(np.median(d)))
    assert float(np.max(d)) <= max_tol, (what, "max", float(np.max(d)))
