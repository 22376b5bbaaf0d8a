#!/bin/bash
# Quick GPU iteration: parity tests + short benches.   gpurun --timeout 900 -- 'bash tools/gpu_iter.sh tag [sanitize]'
TAG=${1:-it}
OUT=gpurun_out/$TAG
mkdir -p $OUT
timeout 600 python -m pytest tests -m gpu -x -q > $OUT/pytest_gpu.log 2>&1; echo "pytest exit $?" >> $OUT/pytest_gpu.log
tail -15 $OUT/pytest_gpu.log
if [ "$2" = "sanitize" ]; then
  timeout 400 compute-sanitizer --tool memcheck --print-limit 5 python -c "import __graft_entry__ as g; g.smoke()" > $OUT/sanitize.log 2>&1
  tail -8 $OUT/sanitize.log
fi
for wl in cfg3 cfg5 cfg2; do
  timeout 300 python bench.py --workload $wl --steps 10 --warmup 3 --no-cpu-baseline --e2e-steps 1 > $OUT/bench_$wl.json 2> $OUT/bench_$wl.err
  echo "bench $wl exit $?"; python - <<PY
import json
try:
    d=json.load(open("$OUT/bench_$wl.json"))
    print("$wl", "ms/step", d["ms_per_step"], "Mtok/s", round(d["value"]/1e6,2), "step_frac", d["roofline"]["step_frac"], [(k["kernel"],k["avg_ms"]) for k in d["kernels"]])
except Exception as e:
    print("no json", e); print(open("$OUT/bench_$wl.err").read()[-1500:])
PY
done
