#!/bin/bash
# Guarded first run of a new kernel: a focused parity subset under a short timeout, then the full iteration script.
TAG=${1:-q}; OUT=gpurun_out/$TAG; mkdir -p $OUT
timeout 150 python -m pytest tests -m gpu -x -q -k "cfg2_shape or selscan_variants or half_precision or strided" > $OUT/quick.log 2>&1
rc=$?; tail -5 $OUT/quick.log; echo "quick exit $rc"
if [ $rc -ne 0 ]; then exit 1; fi
bash tools/gpu_iter.sh $TAG
