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


def test_deposit_block_equals_trace_absorb_profile(lib):
    """gfb_rays_deposit_block (profile resident on the device, no host synchronisation -- the config-3 path
    that is all-reduced with NCCL) accumulates the same profile as gfb_rays_trace_absorb, and that profile
    is the oracle's binning of the per-ray d_power (1e-12)."""
    import torch
    from graph_framework_b200.rays import RayTracer
    from graph_framework_b200 import workloads
    from oracle import port
    n, dt, sub, blocks = 20000, 1.0e-3, 20, 30
    bins, lo, hi = (32, 8, 8), (1.8, -0.5, -0.5), (2.6, 0.5, 0.5)
    state = workloads.efit_ensemble(n, seed=3)
    profiles = []
    for device_resident in (False, True):
        tr = RayTracer("ordinary_wave", "efit", n, dt, options="absorption=1")
        tr.set_state(state)
        tr.init("kx")
        tr.compile()
        if device_resident:
            hist = torch.zeros(bins, dtype=torch.float64, device="cuda")
            oracle = np.zeros(bins)
            for _ in range(blocks):
                tr.deposit_block(sub, hist.data_ptr(), lo, hi, bins)
                pos, ab = tr.get_state(residual=False), tr.get_absorbed()
                oracle += port.deposit(pos["x"], pos["y"], pos["z"], ab["d_power"], lo, hi, bins)
            tr.wait()
            torch.cuda.synchronize()
            profiles.append(hist.cpu().numpy())
            assert np.max(np.abs(profiles[-1] - oracle)) < 1.0e-12*max(np.max(oracle), 1.0)
            assert oracle.sum() > 0.2*n
        else:
            _, _, prof = tr.trace_absorb(blocks, sub, bins=bins, lo=lo, hi=hi, records=False)
            profiles.append(prof)
        tr.close()
    assert np.max(np.abs(profiles[0] - profiles[1])) < 1.0e-12*np.max(profiles[0])


def test_allreduce_sum_over_peer_memory(lib):
    """gfb_allreduce_sum_f64: one buffer per device summed into every buffer (reduce-scatter + all-gather on
    NVLink peer memory), bit-identical on all devices and equal to the sum in device order.  With one GPU the
    call is the identity; the multi-device case needs >= 2 GPUs (gpurun --gpus 2)."""
    import ctypes
    count = lib.gfb_device_count()
    devices = min(count, 8)
    rng = np.random.default_rng(5)
    for n in (1, 7, 4096, 524288 + 3):
        ctxs = [lib.gfb_ctx_create(d) for d in range(devices)]
        data = [rng.normal(size=n)*10.0**rng.integers(-3, 3) for _ in range(devices)]
        for d, (c, a) in enumerate(zip(ctxs, data)):
            assert lib.gfb_buffer(c, 900 + d, a.nbytes, a.ctypes.data_as(ctypes.c_void_p), None) == 0
        arr = (ctypes.c_void_p*devices)(*ctxs)
        keys = (ctypes.c_uint64*devices)(*[900 + d for d in range(devices)])
        for repeat in range(2):                                  # the second call reuses scratch and events
            assert lib.gfb_allreduce_sum_f64(arr, devices, keys, n) == 0, lib.gfb_last_error()
            expect = data[0].copy()
            for a in data[1:]:
                expect = expect + a                              # device order, like the kernel
            outs = []
            for d, c in enumerate(ctxs):
                assert lib.gfb_wait(c) == 0
                out = np.empty(n)
                assert lib.gfb_copy_d2h(c, 900 + d, out.ctypes.data_as(ctypes.c_void_p), out.nbytes) == 0
                outs.append(out)
            for out in outs:
                assert np.array_equal(out, expect), (n, devices, repeat)
            data = [expect.copy() for _ in range(devices)]       # every buffer now holds the sum
        for c in ctxs:
            lib.gfb_ctx_destroy(c)


def test_xrays_driver_reduces_shard_profiles_on_the_devices(lib, tmp_path):
    """xrays with absorption on every visible device: the shards' profiles are summed by
    gfb_allreduce_sum_f64 and equal the binning of all shards' d_power records."""
    from graph_framework_b200 import xrays
    from graph_framework_b200.tools.gfbt import read_gfbt
    devices = min(lib.gfb_device_count(), 4)
    prefix = str(tmp_path / "result")
    rc = xrays.main(["--dispersion=ordinary_wave", "--endtime=0.6", "--equilibrium=efit", "--init_kx",
                     "--init_kx_mean=-700.0", "--init_ky_dist=normal", "--init_ky_mean=-100.0", "--init_ky_sigma=10.0",
                     "--init_kz_dist=normal", "--init_kz_sigma=10.0", "--init_w_dist=normal", "--init_w_mean=700",
                     "--init_w_sigma=10.0", "--init_x_mean=2.5", "--init_y_dist=normal", "--init_y_sigma=0.05",
                     "--init_z_dist=normal", "--init_z_sigma=0.05", "--num_rays=3001", "--num_times=600",
                     "--sub_steps=20", "--use_cyl_xy", "--seed", "--devices=%d" % devices, "--absorption_model=weak_damping",
                     "--num_x=16", "--min_x=1.8", "--max_x=2.6", "--num_y=2", "--min_y=-1", "--max_y=1",
                     "--num_z=2", "--min_z=-1", "--max_z=1", "--output=" + prefix])
    assert rc == 0
    binned = read_gfbt(str(tmp_path / "bins.gfbt"))
    from oracle import port
    total = np.zeros((16, 2, 2))
    rays = 0
    for d in range(devices):
        out = read_gfbt(prefix + "%d.gfbt" % d)
        rays += out["x"].shape[1]
        for b in range(1, out["x"].shape[0]):
            total += port.deposit(out["x"][b], out["y"][b], out["z"][b], out["d_power"][b], (1.8, -1, -1), (2.6, 1, 1), (16, 2, 2))
    assert rays == 3001
    assert np.max(np.abs(binned["bins"]*3001 - total)) < 1.0e-12*np.max(total)
