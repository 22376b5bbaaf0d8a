#!/bin/bash
# bench one workload under several library variants / env settings.  gpurun -- 'bash tools/gpu_variants.sh tag wl "VAR=..." ...'
TAG=$1; WL=$2; shift 2
OUT=gpurun_out/$TAG; mkdir -p $OUT
for cfg in "$@"; do
  name=$(echo "$cfg" | tr ' =' '__')
  env $cfg timeout 300 python bench.py --workload $WL --steps 10 --warmup 3 --no-cpu-baseline --e2e-steps 1 > $OUT/${WL}_$name.json 2> $OUT/${WL}_$name.err
  python - <<PY
import json
try:
    d=json.load(open("$OUT/${WL}_$name.json"))
    print("$WL $cfg | ms/step", d["ms_per_step"], [(k["kernel"],k["avg_ms"]) for k in d["kernels"][:2]])
except Exception as e:
    print("$WL $cfg | no json", e); print(open("$OUT/${WL}_$name.err").read()[-800:])
PY
done
