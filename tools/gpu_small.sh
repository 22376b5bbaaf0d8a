#!/bin/bash
# chained (v2/v4) vs L-split kernels on small-batch shapes.  gpurun -- 'bash tools/gpu_small.sh tag'
TAG=${1:-small}; OUT=gpurun_out/$TAG; mkdir -p $OUT
run() { # name, env, args...
  name=$1; envs=$2; shift 2
  env $envs timeout 300 python bench.py "$@" --steps 10 --warmup 3 --no-cpu-baseline --e2e-steps 1 > $OUT/$name.json 2> $OUT/$name.err
  python - <<PY
import json
try:
    d=json.load(open("$OUT/$name.json"))
    print("$name | ms/step", d["ms_per_step"], [(k["kernel"],k["avg_ms"]) for k in d["kernels"][:4]])
except Exception as e:
    print("$name | no json", e); print(open("$OUT/$name.err").read()[-800:])
PY
}
for b in 1 2 4; do
  run cfg3_b${b}_split GFE_X=0 --workload cfg3 --batch $b
  run cfg3_b${b}_chain GFE_SELSCAN_V2=1 --workload cfg3 --batch $b
done
for wl in prod cfg4 cfg2; do
  run ${wl}_split GFE_X=0 --workload $wl
  run ${wl}_chain GFE_SELSCAN_V2=1 --workload $wl
done
