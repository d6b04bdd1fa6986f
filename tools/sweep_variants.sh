#!/bin/bash
# Sweep launch/emission options of a bench workload (run on the GPU box).
WL=${1:-efit_xmode}
for opt in "unroll_stages=0 minblocks=2" "unroll_stages=0 minblocks=3" "unroll_stages=0 minblocks=4" "unroll_stages=1 minblocks=2" "unroll_stages=1 minblocks=3" "unroll_stages=1 minblocks=4" "unroll_stages=0 minblocks=3 block=64" "unroll_stages=0 minblocks=2 block=256"; do
  echo "== $WL $opt"
  python bench.py --workload $WL --steps 3 --warmup 2 --no-cpu-baseline --options "$opt" 2>&1 | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print('%.3e ray-steps/s  %.2f ms  frac %.3f  regs %d local %d  e2e %.3e'%(d['value'],d['ms_per_step'],d['roofline']['frac'],d['kernel']['registers'],d['kernel']['local_bytes'],d['e2e']['value']))
    elif 'rror' in l: print(l.strip()[:300])
"
done
