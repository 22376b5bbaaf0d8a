#!/bin/bash
# Source-level counters (instructions executed + stall samples per SASS line) of the two chained kernels on cfg3.
TAG=${1:-src1}; OUT=gpurun_out/$TAG; mkdir -p $OUT
timeout 600 ncu --section SourceCounters --section WarpStateStats --import-source on --clock-control none -k "regex:selscan_(fwd_v4|bwd_chain)" -s 4 -c 2 -f -o $OUT/src_cfg3 \
    python bench.py --workload cfg3 --steps 1 --warmup 2 --no-cpu-baseline --e2e-steps 1 > $OUT/run.log 2>&1
ls -la $OUT; tail -3 $OUT/run.log
