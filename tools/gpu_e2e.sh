#!/bin/bash
# e2e leg sweep over the pipeline's largest chunk.  gpurun -- 'bash tools/gpu_e2e.sh tag'
TAG=${1:-e2e}; OUT=gpurun_out/$TAG; mkdir -p $OUT
for rows in 1 2 4 8 16; do
  timeout 300 python bench.py --workload cfg3 --steps 3 --warmup 3 --no-cpu-baseline --e2e-steps 5 --e2e-rows $rows > $OUT/rows$rows.json 2> $OUT/rows$rows.err
  python - <<PY
import json
try:
    d=json.load(open("$OUT/rows$rows.json")); print("rows=$rows | e2e ms", d["e2e"]["ms_per_step"])
except Exception as e:
    print("rows=$rows | no json", e); print(open("$OUT/rows$rows.err").read()[-1200:])
PY
done
