#!/bin/bash
# e2e leg sweep over pipeline chunk sizes + lib variants.  gpurun -- 'bash tools/gpu_e2e.sh tag'
TAG=${1:-e2e}; OUT=gpurun_out/$TAG; mkdir -p $OUT
show() { python - <<PY
import json
try:
    d=json.load(open("$1"))
    print("$2 | ms/step", d["ms_per_step"], [(k["kernel"],k["avg_ms"]) for k in d["kernels"][:2]], "e2e ms", d["e2e"]["ms_per_step"], "rows", d["e2e"].get("rows_per_chunk"))
except Exception as e:
    print("$2 | no json", e); print(open("$1".replace(".json",".err")).read()[-1200:])
PY
}
for rows in 1 2 4 8; do
  timeout 300 python bench.py --workload cfg3 --steps 5 --warmup 3 --no-cpu-baseline --e2e-steps 3 --e2e-rows $rows > $OUT/rows$rows.json 2> $OUT/rows$rows.err
  show $OUT/rows$rows.json "rows=$rows"
done
for wl in cfg3 cfg5; do
  GFE_LIB_VARIANT=mb3 timeout 300 python bench.py --workload $wl --steps 10 --warmup 3 --no-cpu-baseline --e2e-steps 1 > $OUT/mb3_$wl.json 2> $OUT/mb3_$wl.err
  show $OUT/mb3_$wl.json "mb3 $wl"
done
