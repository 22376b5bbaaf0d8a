#!/bin/bash
# A/B of library variants (gfe_mamba_b200/lib/libgfe_mamba_b200.<tag>.so): usage gpu_var.sh OUTTAG "workloads" tag1 tag2 ...  ("-" = the default library)
TAG=$1; WLS=$2; shift 2; OUT=gpurun_out/$TAG; mkdir -p $OUT
for v in "$@"; do
  for w in $WLS; do
    if [ "$v" = "-" ]; then unset GFE_LIB_VARIANT; else export GFE_LIB_VARIANT=$v; fi
    timeout 200 python bench.py --workload $w --steps 12 --warmup 4 --no-cpu-baseline --e2e-steps 1 > $OUT/${w}_$v.json 2> $OUT/${w}_$v.err
    python -c "import json;d=json.load(open('$OUT/${w}_$v.json'));print('$v $w', d['ms_per_step'], [(k['kernel'],k['avg_ms']) for k in d['kernels'][:4]])" || tail -3 $OUT/${w}_$v.err
  done
done
