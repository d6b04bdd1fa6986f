#!/bin/bash
# Run ON AN 8-GPU BOX (gpurun --gpus 8): the default line on 1 GPU, weak and strong scaling of the headline workload,
# config 3 (device deposition + overlapped NCCL all-reduce), the Boris / VMEC totals and the default line at 8 ranks.
# Output: gpurun_out/scale/*.json (last line = the bench line)
set -u
OUT=gpurun_out/scale; mkdir -p $OUT
run() {  # name, ranks, bench args...
  local name=$1 n=$2; shift; shift
  if [ "$n" = 1 ]; then
    timeout 900 python bench.py --gpus 1 "$@" > $OUT/$name.json 2> $OUT/$name.err
  else
    timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29500 + RANDOM % 400)) \
        bench.py --gpus $n "$@" > $OUT/$name.json 2> $OUT/$name.err
  fi
  python - <<PY
import json
try:
    d = json.loads(open("$OUT/$name.json").read().strip().splitlines()[-1])
    e = d.get("e2e") or {}
    print("%-28s n=%d value %.4g ms/step %.3f e2e %.4g %s" % ("$name", d["n_gpus"], d["value"], d["ms_per_step"], e.get("value", 0), d.get("collective", {}).get("share_of_step", "")))
except Exception as ex:
    print("$name FAILED", ex)
PY
}
run default_1gpu 1
for n in 1 2 4 8; do
  run xmode_weak_${n}gpu $n --steps 10 --warmup 3 --no-extras --no-cpu-baseline
  run xmode_strong_${n}gpu $n --steps 10 --warmup 3 --no-extras --no-cpu-baseline --scaling strong
done
for n in 1 2 4 8; do
  run config3_${n}gpu $n --steps 20 --warmup 5 --no-extras --workload efit_absorb
done
run config3_strong_8gpu 8 --steps 20 --warmup 5 --no-extras --workload efit_absorb --scaling strong
run boris_strong_8gpu 8 --steps 3 --warmup 2 --no-extras --no-cpu-baseline --no-e2e --workload boris
run vmec_strong_8gpu 8 --steps 3 --warmup 2 --no-extras --no-cpu-baseline --no-e2e --workload vmec_omode --scaling strong
run default_8gpu 8
