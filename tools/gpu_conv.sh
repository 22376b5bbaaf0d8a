#!/bin/bash
TAG=$1; shift; OUT=gpurun_out/$TAG; mkdir -p $OUT
for v in "$@"; do
  if [ "$v" = "-" ]; then unset GFE_LIB_VARIANT; else export GFE_LIB_VARIANT=$v; fi
  for shape in "16 4096 1536 bf16" "256 1024 1024 bf16" "2 1858 1024 bf16" "16 4096 1536 f32"; do
    python tools/bench_conv1d.py $shape > $OUT/conv_${v}_$(echo $shape | tr ' ' '_').json 2>$OUT/err.txt
    python -c "import json,sys;d=json.load(open('$OUT/conv_${v}_$(echo $shape | tr ' ' '_').json'));print('$v', '$shape', d['ms_per_step'], [(k['kernel'][12:],k['avg_ms'],k['frac_of_measured_peak']) for k in d['kernels']])" || tail -3 $OUT/err.txt
  done
done
