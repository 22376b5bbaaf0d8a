#!/bin/bash
# A/B sweep of environment knobs on one workload.   gpurun -- 'bash tools/gpu_sweep.sh tag cfg3 "A=1 B=2" "A=2" ...'
TAG=$1; WL=$2; shift 2
OUT=gpurun_out/$TAG; mkdir -p $OUT
for cfg in "$@"; do
  name=$(echo "$cfg" | tr ' =' '__')
  env $cfg timeout 300 python bench.py --workload $WL --steps 10 --warmup 3 --no-cpu-baseline --e2e-steps 1 > $OUT/$name.json 2> $OUT/$name.err
  python - <<PY
import json
try:
    d=json.load(open("$OUT/$name.json"))
    print("$cfg |", "ms/step", d["ms_per_step"], [(k["kernel"],k["avg_ms"]) for k in d["kernels"][:2]])
except Exception as e:
    print("$cfg | no json", e); print(open("$OUT/$name.err").read()[-800:])
PY
done
