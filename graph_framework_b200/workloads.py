"""Synthetic ray ensembles for the BASELINE.json configurations.

Inputs are generated once on the host with a seeded numpy generator and the same
arrays are given to the GPU path and to the oracle, so parity never depends on
reproducing a C++ random number stream.
"""
import numpy as np

#  Algorithmic FP64 flops per unit, counted from the reference's emitted kernels
#  (add/sub/mul/div/sqrt/pow = 1, fma = 2).
FLOP_PER_RAY_STEP = {
    ("cold_plasma", "efit"): 4392.0,
    ("ordinary_wave", "efit"): 2131.0,
    ("extra_ordinary_wave", "efit"): 2324.0,
    ("cold_plasma", "slab"): 856.0,
    ("cold_plasma", "slab_density"): 612.0,
    ("ordinary_wave", "slab_density"): 96.0,
    ("cold_plasma", "vmec"): 63165.0,
}
FLOP_PER_PARTICLE_STEP_BORIS = 216.0
#  HBM bytes per ray-step when state makes one round trip per launch of S fused steps.
STATE_BYTES_PER_RAY = 8*8 + 7*8 + 8      # 8 loads, 7 stores, residual


def bench_rays(n):
    """C1: xrays_bench initial conditions, every ray identical (xrays_bench.cpp:62-71)."""
    s = {k: np.zeros(n) for k in ("t", "w", "x", "y", "z", "kx", "ky", "kz")}
    s["w"][:] = 500.0
    s["x"][:] = 2.5
    s["kx"][:] = -600.0
    return s


def efit_ensemble(n, seed=0, radius=2.5):
    """C2: the distribution of graph_driver/efit_example.sh (xrays.cpp:448-453 draw order
    w, kx, ky, kz, z, then cylindrical x, y with radius 2.5 and angle ~ N(0, 0.05)).
    `radius` moves the launch circle (the absorption workload starts just outside the resonance)."""
    rng = np.random.default_rng(seed)
    s = {"t": np.zeros(n)}
    s["w"] = rng.normal(700.0, 10.0, n)
    s["kx"] = np.full(n, -700.0)
    s["ky"] = rng.normal(-100.0, 10.0, n)
    s["kz"] = rng.normal(0.0, 10.0, n)
    s["z"] = rng.normal(0.0, 0.05, n)
    phi = rng.normal(0.0, 0.05, n)
    s["x"] = radius*np.cos(phi)
    s["y"] = radius*np.sin(phi)
    return s


def slab_ensemble(n, seed=0):
    """Analytic O-mode variant: w = 1100, x ~ U(-0.5, 0.5), k = (1000, ky, kz)."""
    rng = np.random.default_rng(seed)
    s = {"t": np.zeros(n)}
    s["w"] = np.full(n, 1100.0)
    s["x"] = rng.uniform(-0.5, 0.5, n)
    s["y"] = rng.normal(0.0, 0.05, n)
    s["z"] = rng.normal(0.0, 0.05, n)
    s["kx"] = np.full(n, 1000.0)
    s["ky"] = rng.normal(0.0, 10.0, n)
    s["kz"] = rng.normal(0.0, 10.0, n)
    return s


def interior_states(n, seed=1):
    """Generic states inside the EFIT plasma, for right-hand-side parity checks."""
    rng = np.random.default_rng(seed)
    s = {"t": np.zeros(n)}
    s["w"] = 500.0 + rng.normal(0.0, 5.0, n)
    phi = rng.normal(0.0, 0.3, n)
    r = rng.uniform(1.2, 2.2, n)
    s["x"] = r*np.cos(phi)
    s["y"] = r*np.sin(phi)
    s["z"] = rng.normal(0.0, 0.3, n)
    s["kx"] = -400.0 + rng.normal(0.0, 20.0, n)
    s["ky"] = rng.normal(0.0, 50.0, n)
    s["kz"] = rng.normal(0.0, 50.0, n)
    return s


def vmec_states(n, seed=0):
    """C4: flux coordinates (s, u, v) stored in the x, y, z slots: s ~ U(0.2, 0.8), u, v ~ U(0, 2 pi),
    w = 1000 (above the central plasma frequency of the 1e19 m^-3 profile), covariant wave numbers
    k_u, k_v ~ N(0, 20) and a radial guess k_s = 100 for the Newton solve.  No reference run exists
    for this configuration."""
    rng = np.random.default_rng(seed)
    s = {"t": np.zeros(n), "w": np.full(n, 1000.0)}
    s["x"] = rng.uniform(0.2, 0.8, n)
    s["y"] = rng.uniform(0.0, 2.0*np.pi, n)
    s["z"] = rng.uniform(0.0, 2.0*np.pi, n)
    s["kx"] = np.full(n, 100.0)
    s["ky"] = rng.normal(0.0, 20.0, n)
    s["kz"] = rng.normal(0.0, 20.0, n)
    return s


def boris_ensemble(n, seed=0):
    """C5: xkorc start (xkorc.cpp:47-64) jittered so particles do not share one table cell."""
    rng = np.random.default_rng(seed)
    x = rng.normal(1.7, 0.05, n)
    y = np.zeros(n)
    z = rng.normal(0.0, 0.05, n)
    speed = 0.9951
    pitch = rng.uniform(-0.5, 0.5, n)
    gyro = rng.uniform(0.0, 2.0*np.pi, n)
    upar = speed*np.sin(pitch)
    uperp = speed*np.cos(pitch)
    ux = uperp*np.cos(gyro)*0.1
    uz = uperp*np.sin(gyro)*0.1
    uy = np.sqrt(np.maximum(speed**2 - ux**2 - uz**2, 0.0))*np.sign(np.cos(pitch))
    del upar
    return x, y, z, ux, uy, uz


ORDER = ("t", "w", "x", "y", "z", "kx", "ky", "kz")


def pack(state):
    """8 x N array in the order t, w, x, y, z, kx, ky, kz."""
    return np.stack([np.asarray(state[k], dtype=np.float64) for k in ORDER])


def unpack(arr):
    return {k: np.array(arr[i]) for i, k in enumerate(ORDER)}
