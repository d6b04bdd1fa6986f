"""Largest relative deviation of the VMEC right-hand side from the reference's golden vectors (GPU box)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from graph_framework_b200.rays import RayTracer
ORDER = ("t", "w", "x", "y", "z", "kx", "ky", "kz")
RHS = ("dxdt", "dydt", "dzdt", "dkxdt", "dkydt", "dkzdt", "D")
for opts in ("mode_recurrence=0", "mode_recurrence=1"):
    g = np.load("tests/golden/ref_rhs_ordinary_wave_vmec.npz")
    state = {k: np.array(g["state"][i]) for i, k in enumerate(ORDER)}
    tr = RayTracer("ordinary_wave", "vmec", state["w"].size, 1.0e-4, options=opts)
    tr.set_state(state)
    got = tr.rhs()
    tr.close()
    out = []
    for i, k in enumerate(RHS):
        ref = g["rhs"][i]
        scale = np.maximum(np.abs(ref), 1e-9*np.max(np.abs(ref)) + 1e-300)
        out.append("%s %.1e" % (k, np.max(np.abs(got[k] - ref)/scale)))
    print(opts, " ".join(out))
