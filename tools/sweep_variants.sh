#!/bin/bash
# Sweep launch/emission options of the headline workload (run on the GPU box).
WL=${1:-efit_xmode}
for opt in "minblocks=1" "minblocks=3" "minblocks=4" "minblocks=5" "block=256 minblocks=2" "block=64 minblocks=8" "unroll_stages=1 minblocks=4" "fast_div=0 minblocks=4" "stage_tables=0 minblocks=4"; do
  echo "== $WL $opt"
  python bench.py --workload $WL --steps 3 --warmup 2 --no-cpu-baseline --options "$opt" 2>&1 | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print('%.3e ray-steps/s  %.2f ms  frac %.3f  regs %d local %d  e2e %.3e'%(d['value'],d['ms_per_step'],d['roofline']['frac'],d['kernel']['registers'],d['kernel']['local_bytes'],d['e2e']['value']))
    elif 'rror' in l: print(l.strip()[:300])
"
done
