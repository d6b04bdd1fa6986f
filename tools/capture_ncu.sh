#!/bin/bash
# Run ON THE GPU BOX (under gpurun): plain bench run, then ONE `ncu --set full` capture of the workload's
# step kernel inside the same command, then the JSON bench.py reads.  usage: tools/capture_ncu.sh <workload> [round] [extra bench args]
#   -> gpurun_out/ncu/<workload>.{json,ncu-rep,raw.csv,plain.json}   (copy the .json to profiles/r<round>_ncu_<workload>.json)
set -u
W=$1; R=${2:-2}; shift; shift || true
OUT=gpurun_out/ncu; mkdir -p $OUT
KERNEL=solver_kernel; UNITS=1000000
case $W in
  boris) KERNEL=step; UNITS=20000000; ARGS="--workload boris --rays 20000000 --scaling weak";;
  vmec_*) UNITS=1250000; ARGS="--workload $W --options bin_rays=100";;     # order checked every 100 steps: one launch = one bench step
  *) ARGS="--workload $W";;
esac
CMD="python bench.py $ARGS --steps 2 --warmup 3 --no-extras --no-cpu-baseline --no-e2e $*"
GIT_SHA=$(cat .git_sha 2>/dev/null || echo unknown)
$CMD > $OUT/$W.plain.json 2> $OUT/$W.plain.err || { echo "plain run failed"; tail -5 $OUT/$W.plain.err; exit 1; }
# the kernel regex is anchored: "step" must not match solver_kernel's helpers
ncu --set full --clock-control none --import-source on -k regex:"^$KERNEL\$" -s 3 -c 1 -f -o $OUT/$W $CMD > $OUT/$W.ncu.log 2>&1 || { echo "ncu failed"; tail -5 $OUT/$W.ncu.log; exit 1; }
ncu -i $OUT/$W.ncu-rep --page raw --csv > $OUT/$W.raw.csv
python tools/ncu_to_json.py $OUT/$W.raw.csv --workload $W --kernel $KERNEL --units $UNITS --unit-steps 100 \
    --bench-line $OUT/$W.plain.json --command "$CMD" --out $OUT/$W.json
python - <<PY
import json
p="$OUT/$W.json"; d=json.load(open(p)); d["git_sha"]="$GIT_SHA"; json.dump(d, open(p,"w"), indent=1)
PY
