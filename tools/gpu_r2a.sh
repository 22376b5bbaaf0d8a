#!/bin/bash
# Round 2, first run: parity of the new chained kernels (channel-block widths 16 / 32 / 64) and a bench sweep over them.
TAG=${1:-r2a}; OUT=gpurun_out/$TAG; mkdir -p $OUT
K="chained or cfg3 or full_size or half_precision or golden_reference or block_golden"
timeout 500 python -m pytest tests -m gpu -x -q -k "$K" > $OUT/parity_default.log 2>&1; echo "default exit $?" | tee -a $OUT/parity_default.log
tail -4 $OUT/parity_default.log
for c in 16 32; do
  GFE_LIB_VARIANT=exp GFE_FWD_CPC=$c GFE_BWD_CPC=$c timeout 400 python -m pytest tests -m gpu -x -q -k "chained or half_precision" > $OUT/parity_cpc$c.log 2>&1
  echo "cpc$c exit $?" | tee -a $OUT/parity_cpc$c.log; tail -3 $OUT/parity_cpc$c.log
done
bash tools/gpu_variants.sh $TAG cfg3 "GFE_LIB_VARIANT=exp GFE_FWD_CPC=64 GFE_BWD_CPC=64" "GFE_LIB_VARIANT=exp GFE_FWD_CPC=32 GFE_BWD_CPC=32" "GFE_LIB_VARIANT=exp GFE_FWD_CPC=16 GFE_BWD_CPC=16" "GFE_LIB_VARIANT=exp2 GFE_FWD_CPC=64 GFE_BWD_CPC=64" "GFE_LIB_VARIANT=exp2 GFE_FWD_CPC=32 GFE_BWD_CPC=32"
bash tools/gpu_variants.sh $TAG cfg5 "GFE_LIB_VARIANT=exp GFE_FWD_CPC=64 GFE_BWD_CPC=64" "GFE_LIB_VARIANT=exp GFE_FWD_CPC=32 GFE_BWD_CPC=32" "GFE_LIB_VARIANT=exp GFE_FWD_CPC=16 GFE_BWD_CPC=16" "GFE_LIB_VARIANT=exp2 GFE_FWD_CPC=32 GFE_BWD_CPC=32"
