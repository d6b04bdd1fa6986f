"""Step-by-step device-layer diagnostic (run on the GPU box with GFB_DEBUG=1)."""
import faulthandler
import sys
import os
faulthandler.enable()
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np
from graph_framework_b200.graph import Context

g = Context()
n = 8
x = g.variable(n, "x", np.arange(n, dtype=float))
y = g.variable(n, "y", np.ones(n))
out = x*y + 2.0
g.add_item([x, y], [out], [(x + 1.0, x)], "diag", n)
print("compile", flush=True)
g.compile()
print("run", flush=True)
g.run()
print("copy", flush=True)
print(g.copy_to_host(out, n), g.copy_to_host(x, n), flush=True)
g.close()
print("graph ok", flush=True)
from graph_framework_b200.rays import RayTracer
from graph_framework_b200 import workloads
s = workloads.slab_ensemble(n, seed=1)
tr = RayTracer("simple", "slab", n, 1e-2)
tr.set_state(s)
print("rhs", flush=True)
print(tr.rhs()["dxdt"], flush=True)
tr.init("")
tr.compile()
tr.step(3)
print(tr.get_state()["x"], flush=True)
tr.close()
print("rays ok", flush=True)
