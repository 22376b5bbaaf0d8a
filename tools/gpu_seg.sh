#!/bin/bash
# Independent-segment path: focused parity subset first, then the whole GPU suite and the shapes it serves.
TAG=${1:-seg1}; OUT=gpurun_out/$TAG; mkdir -p $OUT
timeout 300 python -m pytest tests -m gpu -x -q -k "l_split or independent or strided or fp32_vs_oracle or variants" > $OUT/quick.log 2>&1
rc=$?; tail -15 $OUT/quick.log; echo "quick exit $rc"
if [ $rc -ne 0 ]; then exit 1; fi
for w in prod cfg4 cfg3; do
timeout 300 python bench.py --workload $w --steps 20 --warmup 5 --no-cpu-baseline --e2e-steps 2 > $OUT/bench_$w.json 2> $OUT/bench_$w.err
python -c "import json;d=json.load(open('$OUT/bench_$w.json'));print('$w', d['ms_per_step'], [(k['kernel'],k['avg_ms']) for k in d['kernels']])" || tail -5 $OUT/bench_$w.err
done
timeout 1200 python -m pytest tests -m gpu -q > $OUT/pytest_gpu.log 2>&1; echo "pytest exit $?" | tee -a $OUT/pytest_gpu.log
tail -8 $OUT/pytest_gpu.log
