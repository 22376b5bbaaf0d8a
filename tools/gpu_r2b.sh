#!/bin/bash
TAG=${1:-r2b}; OUT=gpurun_out/$TAG; mkdir -p $OUT
timeout 500 python -m pytest tests -m gpu -x -q -k "chained or cfg3 or full_size or half_precision or golden_reference or block_golden or optim or step or inner" > $OUT/parity.log 2>&1; echo "parity exit $?" | tee -a $OUT/parity.log
tail -6 $OUT/parity.log
bash tools/gpu_variants.sh $TAG cfg3 "GFE_LIB_VARIANT=" "GFE_LIB_VARIANT=nokeep" "GFE_LIB_VARIANT=exp GFE_BWD_CPC=32"
bash tools/gpu_variants.sh $TAG cfg5 "GFE_LIB_VARIANT=" "GFE_LIB_VARIANT=nokeep" "GFE_LIB_VARIANT=exp GFE_BWD_CPC=32"
