#!/bin/bash
# ncu full capture of the scan kernels on one workload.  gpurun --timeout 900 -- 'bash tools/gpu_prof.sh tag [workload] [regex]'
TAG=${1:-prof}; WL=${2:-cfg3}; RX=${3:-selscan_(fwd|bwd)_v3}; SKIP=${4:-6}; CNT=${5:-2}
OUT=gpurun_out/$TAG
mkdir -p $OUT
timeout 900 ncu --set full --clock-control none --import-source on -k "regex:$RX" -s $SKIP -c $CNT -f -o $OUT/prof_$WL \
    python bench.py --workload $WL --steps 1 --warmup 3 --no-cpu-baseline --e2e-steps 1 > $OUT/prof_run.log 2>&1
tail -3 $OUT/prof_run.log
ls -la $OUT
