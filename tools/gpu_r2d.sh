#!/bin/bash
TAG=${1:-r2d}; OUT=gpurun_out/$TAG; mkdir -p $OUT
timeout 900 python -m pytest tests -m gpu -q > $OUT/pytest_gpu.log 2>&1; echo "pytest exit $?" | tee -a $OUT/pytest_gpu.log
tail -8 $OUT/pytest_gpu.log
bash tools/gpu_variants.sh $TAG cfg3 "GFE_LIB_VARIANT=exp GFE_CHAIN_NSEG=16" "GFE_LIB_VARIANT=exp GFE_CHAIN_NSEG=32"
for b in 4 8 12 16 24 32 48; do
  timeout 200 python bench.py --workload cfg3 --batch $b --steps 8 --warmup 3 --no-cpu-baseline --e2e-steps 1 > $OUT/b$b.json 2> $OUT/b$b.err
  python - <<PY
import json
try:
    d=json.load(open("$OUT/b$b.json")); print("cfg3 B=$b ms/step", d["ms_per_step"], "us per row", round(1e3*d["ms_per_step"]/$b,2), [(k["kernel"],k["avg_ms"]) for k in d["kernels"][:2]])
except Exception as e: print("B=$b no json", e)
PY
done
