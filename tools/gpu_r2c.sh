#!/bin/bash
TAG=${1:-r2c}; OUT=gpurun_out/$TAG; mkdir -p $OUT
timeout 900 python -m pytest tests -m gpu -x -q > $OUT/pytest_gpu.log 2>&1; echo "pytest exit $?" | tee -a $OUT/pytest_gpu.log
tail -6 $OUT/pytest_gpu.log
bash tools/gpu_variants.sh $TAG cfg3 "GFE_LIB_VARIANT=" "GFE_LIB_VARIANT=exp GFE_CHAIN_NSEG=10" "GFE_LIB_VARIANT=exp GFE_CHAIN_NSEG=4"
bash tools/gpu_variants.sh $TAG cfg5 "GFE_LIB_VARIANT="
bash tools/gpu_variants.sh $TAG cfg2 "GFE_LIB_VARIANT="
