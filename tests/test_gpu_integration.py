"""The drop-in boundary exercised by the UNMODIFIED reference front end (GPU).

integration/_build/ref_driver_b200 is oracle/ref_driver.cpp -- the reference's own graph,
reductions, autodiff, solver::rk4 and dispersion_interface::solve -- compiled in the build
container against integration/b200_context.hpp + libgfb200.so instead of gpu::cuda_context.
Its results on the GPU must equal the reference CPU path's golden vectors: same expressions,
different device layer."""
import os
import subprocess
import tempfile

import numpy as np
import pytest

from conftest import ROOT, golden, rel_dev

pytestmark = pytest.mark.gpu

DRIVER = os.path.join(ROOT, "integration", "_build", "ref_driver_b200")


def run_trace(disp, eq, solver, state, dt, nsteps, every, init):
    n = state.shape[1]
    with tempfile.TemporaryDirectory() as d:
        fin, fout = os.path.join(d, "in.bin"), os.path.join(d, "out.bin")
        np.ascontiguousarray(state).tofile(fin)
        env = dict(os.environ, GFB_EFIT_FILE=os.path.join(ROOT, "tests", "golden", "efit.gfbt"))
        subprocess.run([DRIVER, "trace", disp, eq, solver, str(n), repr(float(dt)), str(nsteps), str(every), init, fin, fout],
                       check=True, cwd=ROOT, env=env, timeout=900)
        return np.fromfile(fout).reshape(-1, 9, n)


@pytest.mark.skipif(not os.path.exists(DRIVER), reason="integration/_build not built (needs the reference tree)")
@pytest.mark.parametrize("disp,eq,solver", [("extra_ordinary_wave", "efit", "rk4"), ("cold_plasma", "slab_density", "rk4")])
def test_reference_front_end_on_b200_context(disp, eq, solver):
    g = golden("ref_trace_%s_%s_%s" % (disp, eq, solver))
    ref = g["per_step"]
    got = run_trace(disp, eq, solver, g["state"], float(g["dt"]), 5, 1, "kx")
    assert got.shape == ref.shape
    # record 0: Newton root through converge_item on gfb_max; records 1..5: fused/deferred steps
    for rec in range(ref.shape[0]):
        for i in range(8):
            assert rel_dev(got[rec][i], ref[rec][i]) < 1.0e-11, (rec, i, rel_dev(got[rec][i], ref[rec][i]))


C_BINDING_TEST = os.path.join(ROOT, "integration", "_build", "c_binding_reference_test")


@pytest.mark.skipif(not os.path.exists(C_BINDING_TEST), reason="integration/_build not built (needs the reference tree)")
def test_reference_c_binding_test_passes_on_libgfb200():
    """The reference's own graph_tests/c_binding_test.c with its own header, unmodified, linked against
    libgfb200.so: run_tests(DOUBLE, false) asserts node identity after reduction, the line
    example and its four derivatives, the converge item, piecewise_1D/2D and index_1D/2D.  (The off-path
    random/complex entry points are supplied by integration/c_binding_reference_test.c, see there.)"""
    out = subprocess.run([C_BINDING_TEST], cwd=ROOT, capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stdout + out.stderr
    assert "all assertions passed" in out.stdout
