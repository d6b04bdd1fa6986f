"""Which records deviate from the golden absorption vectors? (GPU box)"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np
from graph_framework_b200.rays import RayTracer
from graph_framework_b200 import workloads
from test_emit_cpu import absorption_zeta
g = dict(np.load("tests/golden/ref_absorb_ordinary_wave_efit.npz"))
n = g["state"].shape[1]
tr = RayTracer("ordinary_wave", "efit", n, float(g["dt"]), options="absorption=1 " + (sys.argv[1] if len(sys.argv) > 1 else ""))
tr.set_state(workloads.unpack(g["state"]))
tr.init("kx"); tr.compile()
rec, absorbed, _ = tr.trace_absorb(g["records"].shape[0] - 1, int(g["sub_steps"]))
zeta = absorption_zeta(rec)
dev = np.abs(absorbed[:, 0] - g["kamp_im"][1:])
for j, i in np.argwhere(dev > 1e-11 + 1e-7*np.abs(g["kamp_im"][1:]))[:20]:
    print(j, i, "zeta %.4f" % zeta[j, i], "gpu", absorbed[j, 0, i], "ref", g["kamp_im"][1 + j, i])
print("power dev", np.abs(absorbed[:, 1] - g["power"][1:]).max())
