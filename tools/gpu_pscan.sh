#!/bin/bash
TAG=$1; shift; OUT=gpurun_out/$TAG; mkdir -p $OUT
timeout 300 python -m pytest tests -m gpu -x -q -k "pscan or other_state_size or block_reference_goldens" 2>&1 | tail -3
for v in "$@"; do
  if [ "$v" = "-" ]; then unset GFE_LIB_VARIANT; else export GFE_LIB_VARIANT=$v; fi
  for shape in "8 1024 1024 16" "2 1858 1024 16" "32 256 512 16" "1 4096 256 16"; do
    echo -n "$v $shape: "; timeout 120 python tools/bench_pscan.py $shape 2>&1 | tail -1 | python -c "import json,sys;d=json.loads(sys.stdin.read());print(d['ms_per_step'], [(k['kernel'],k['avg_ms'],k['frac_of_measured_peak']) for k in d['kernels']])"
  done
done
