#!/bin/bash
# A/B of the v2 / v3 scan kernels + ncu full captures of both.  gpurun --timeout 1200 -- 'bash tools/gpu_ab.sh tag'
TAG=${1:-ab}
OUT=gpurun_out/$TAG
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $OUT/gpu.txt 2>&1
timeout 600 python -m pytest tests -m gpu -x -q > $OUT/pytest_gpu.log 2>&1; echo "pytest exit $?" >> $OUT/pytest_gpu.log
tail -3 $OUT/pytest_gpu.log
for v in v3 v2; do
  if [ $v = v2 ]; then export GFE_SELSCAN_V3=0; fi
  timeout 300 python bench.py --workload cfg3 --steps 10 --warmup 3 --no-cpu-baseline --e2e-steps 1 > $OUT/bench_cfg3_$v.json 2> $OUT/bench_cfg3_$v.err
  python - <<PY
import json
try:
    d=json.load(open("$OUT/bench_cfg3_$v.json"))
    print("$v", "ms/step", d["ms_per_step"], [(k["kernel"],k["avg_ms"]) for k in d["kernels"]])
except Exception as e:
    print("no json", e); print(open("$OUT/bench_cfg3_$v.err").read()[-1500:])
PY
  timeout 600 ncu --set full --clock-control none --import-source on -k "regex:selscan_(fwd|bwd)_$v" -s 6 -c 2 -f -o $OUT/prof_$v \
      python bench.py --workload cfg3 --steps 1 --warmup 3 --no-cpu-baseline --e2e-steps 1 > $OUT/prof_run_$v.log 2>&1
done
ls -la $OUT
