#!/bin/bash
# Occupancy variants of the step kernel: block size x promised blocks/SM (run on the GPU box).
WL=${1:-efit_xmode}
for opt in "" "block=64 minblocks=6" "block=64 minblocks=7" "block=96 minblocks=4" "block=96 minblocks=5" "block=160 minblocks=2" "block=192 minblocks=2" "block=32 minblocks=12" "block=32 minblocks=14"; do
  echo "== $WL [$opt]"
  python bench.py --workload $WL --steps 3 --warmup 2 --no-cpu-baseline --options "$opt" 2>&1 | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print('%.3e ray-steps/s  %.2f ms  frac %.3f  regs %d local %d  e2e %.3e'%(d['value'],d['ms_per_step'],d['roofline']['frac'],d['kernel']['registers'],d['kernel']['local_bytes'],d['e2e']['value']))
    elif 'rror' in l: print(l.strip()[:300])
"
done
