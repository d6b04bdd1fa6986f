#!/bin/bash
# A/B of the text passes on the reference front end + b200_context (GPU box): 6 timed blocks each.
python - <<PY
import sys
sys.path.insert(0, ".")
from graph_framework_b200 import workloads
workloads.pack(workloads.efit_ensemble(1000000, seed=0)).tofile("/tmp/arm_in.bin")
PY
for d in extra_ordinary_wave cold_plasma; do
  for mode in all divides_only; do
    if [ $mode = divides_only ]; then export GFB_B200_DIVIDES_ONLY=1; else unset GFB_B200_DIVIDES_ONLY; fi
    GFB_EFIT_FILE=tests/golden/efit.gfbt integration/_build/ref_driver_b200 bench $d efit 1000000 2e-5 100 1 /tmp/arm_in.bin 6 2>/dev/null | tail -1 | python -c "
import sys, json
d = json.loads(sys.stdin.read()); b = sorted(d['block_s']); print('$d $mode median block %.3f ms -> %.3e ray-steps/s' % (1e3*b[len(b)//2], 1e8/b[len(b)//2]))"
  done
done
