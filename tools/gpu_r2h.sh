#!/bin/bash
TAG=${1:-r2h}; OUT=gpurun_out/$TAG; mkdir -p $OUT
timeout 900 python -m pytest tests -m gpu -q -s > $OUT/pytest_gpu.log 2>&1; echo "pytest exit $?" | tee -a $OUT/pytest_gpu.log
grep -E "^cfg3|^cfg4|^prod|^cfg5" $OUT/pytest_gpu.log; tail -4 $OUT/pytest_gpu.log
python tools/bench_pscan.py 8 1024 1024 16 > $OUT/pscan.txt 2>&1; python tools/bench_pscan.py 2 1858 1024 16 >> $OUT/pscan.txt 2>&1; python tools/bench_pscan.py 32 256 512 16 >> $OUT/pscan.txt 2>&1; cat $OUT/pscan.txt
for r in 2 4 8; do
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --e2e-rows $r > $OUT/cfg3_rows$r.json 2> $OUT/cfg3_rows$r.err
python -c "import json;d=json.load(open('$OUT/cfg3_rows$r.json'));print('rows $r', d['ms_per_step'], [(k['kernel'],k['avg_ms']) for k in d['kernels'][:3]], 'e2e', d['e2e']['ms_per_step'])"
done
