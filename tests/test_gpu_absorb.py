"""Power absorption along traced rays on the GPU against the reference's own
kernels (golden vectors from oracle/_ref: absorption::weak_damping in complex<double> with
SAFE_MATH, then the bin_power stage) and against the numpy oracle."""
import os
import sys

import numpy as np
import pytest

from conftest import ROOT, golden, rel_dev

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def traced(lib):
    from graph_framework_b200.rays import RayTracer
    from graph_framework_b200 import workloads
    g = golden("ref_absorb_ordinary_wave_efit")
    n = g["state"].shape[1]
    blocks = g["records"].shape[0] - 1
    tr = RayTracer("ordinary_wave", "efit", n, float(g["dt"]), options="absorption=1")
    tr.set_state(workloads.unpack(g["state"]))
    tr.init("kx")
    tr.compile()
    lo, hi, bins = (1.9, -0.4, -0.3), (2.6, 0.4, 0.3), (14, 8, 6)
    rec, absorbed, profile = tr.trace_absorb(blocks, int(g["sub_steps"]), bins=bins, lo=lo, hi=hi)
    yield g, rec, absorbed, profile, (lo, hi, bins)
    tr.close()


def test_trajectory_matches_reference(traced):
    g, rec, absorbed, profile, _ = traced
    for i in range(8):
        assert rel_dev(rec[:, i], g["records"][1:, i]) < 1.0e-10, i


def test_damping_and_power_match_reference(traced):
    """Im k_amp: relative 1e-7 + absolute 1e-11 (it carries exp(-zeta^2), zeta^2 up to 740, evaluated
    on a trajectory that itself agrees to ~1e-12); power and d_power: absolute 1e-10."""
    g, rec, absorbed, profile, _ = traced
    kamp, power, d_power = absorbed[:, 0], absorbed[:, 1], absorbed[:, 2]
    assert np.isfinite(absorbed).all()
    assert np.max(np.abs(kamp - g["kamp_im"][1:]) - 1.0e-7*np.abs(g["kamp_im"][1:])) < 1.0e-11
    assert np.max(kamp) > 5.0
    assert np.max(np.abs(power - g["power"][1:])) < 1.0e-10
    assert np.max(np.abs(d_power - g["d_power"][1:])) < 1.0e-10
    assert 0.05 < power[-1].min() < 0.6


def test_profile_matches_oracle_binning(traced):
    sys.path.insert(0, ROOT)
    from oracle import port
    g, rec, absorbed, profile, (lo, hi, bins) = traced
    expect = np.zeros(bins)
    for b in range(rec.shape[0]):
        expect += port.deposit(rec[b, 2], rec[b, 3], rec[b, 4], absorbed[b, 2], lo, hi, bins)
    assert expect.sum() > 1.0
    assert np.max(np.abs(profile - expect)) < 1.0e-12*expect.max()


def test_absorption_against_port_on_a_larger_ensemble(lib, efit_tables):
    """4096 rays of the efit_example distribution: damping rate and transmitted power against the
    numpy oracle evaluated on the GPU's own records; the profile holds exactly the binned d_power."""
    sys.path.insert(0, ROOT)
    from oracle import port
    from graph_framework_b200.rays import RayTracer
    from graph_framework_b200 import workloads
    n, blocks, sub = 4096, 12, 50
    tr = RayTracer("ordinary_wave", "efit", n, 1.0e-3, options="absorption=1")
    tr.set_state(workloads.efit_ensemble(n, seed=3))
    tr.init("kx")
    tr.compile()
    first = tr.get_state()
    lo, hi, bins = (1.0, -1.0, -1.0), (2.6, 1.0, 1.0), (32, 4, 4)
    rec, absorbed, profile = tr.trace_absorb(blocks, sub, bins=bins, lo=lo, hi=hi)
    tr.close()
    eq = port.make_equilibrium("efit", efit_tables)
    kamp = np.array([port.weak_damping(eq, dict(zip(port.ORDER, r[:8]))).imag for r in rec])
    assert np.max(np.abs(absorbed[:, 0] - kamp) - 1.0e-7*np.abs(kamp)) < 1.0e-10
    xyz = np.concatenate([np.stack([first["x"], first["y"], first["z"]])[None], rec[:, 2:5]])
    power, d_power = port.power_stage(xyz, np.concatenate([np.zeros((1, n)), absorbed[:, 0]]))
    assert np.max(np.abs(absorbed[:, 1] - power[1:])) < 1.0e-12
    assert np.max(np.abs(absorbed[:, 2] - d_power[1:])) < 1.0e-12
    assert np.isfinite(absorbed).all() and np.all(absorbed[:, 1] > 0.0)
    inside = np.ones_like(absorbed[:, 2], dtype=bool)
    for a, (l, h) in zip((2, 3, 4), zip(lo, hi)):
        inside &= (rec[:, a] >= l) & (rec[:, a] < h)
    assert abs(profile.sum() - absorbed[:, 2][inside].sum()) < 1.0e-9*max(profile.sum(), 1.0)
    assert profile.sum() > 0.2*n                                       # the beam is absorbed


def test_xrays_driver_with_absorption(lib, tmp_path):
    from graph_framework_b200 import xrays
    from graph_framework_b200.tools.gfbt import read_gfbt
    prefix = str(tmp_path / "result")
    rc = xrays.main(["--dispersion=ordinary_wave", "--endtime=0.6", "--equilibrium=efit", "--init_kx",
                     "--init_kx_mean=-700.0", "--init_ky_dist=normal", "--init_ky_mean=-100.0", "--init_ky_sigma=10.0",
                     "--init_kz_dist=normal", "--init_kz_sigma=10.0", "--init_w_dist=normal", "--init_w_mean=700",
                     "--init_w_sigma=10.0", "--init_x_mean=2.5", "--init_y_dist=normal", "--init_y_sigma=0.05",
                     "--init_z_dist=normal", "--init_z_sigma=0.05", "--num_rays=2000", "--num_times=600",
                     "--sub_steps=20", "--use_cyl_xy", "--seed", "--devices=1", "--absorption_model=weak_damping",
                     "--num_x=16", "--min_x=1.8", "--max_x=2.6", "--num_y=1", "--min_y=-1", "--max_y=1",
                     "--num_z=1", "--min_z=-1", "--max_z=1", "--output=" + prefix])
    assert rc == 0
    out = read_gfbt(prefix + "0.gfbt")
    assert out["power"].shape == (31, 2000) and out["d_power"].shape == (31, 2000) and out["kamp"].shape == (31, 2000)
    assert np.all(out["power"][0] == 1.0) and np.all(out["d_power"][0] == 0.0)       # record 0: before the first step
    binned = read_gfbt(str(tmp_path / "bins.gfbt"))
    assert binned["bins"].shape == (16, 1, 1) and binned["xbins"].shape == (17,)
    assert abs(binned["bins"].sum()*2000 - out["d_power"].sum()) < 1.0e-6*out["d_power"].sum()
    peak = binned["xbins"][np.argmax(binned["bins"][:, 0, 0])]
    assert 1.95 < peak < 2.25                                          # deposition at the resonance layer
