#!/bin/bash
# One gpurun call: GPU parity tests, smoke, bench (cfg3 + others), ncu launch list and full captures of the two scan kernels.
#   /usr/local/graft/bin/gpurun --timeout 1500 -- 'bash tools/gpu_check.sh [tag]'
TAG=${1:-r01}
OUT=gpurun_out/$TAG
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $OUT/gpu.txt 2>&1

timeout 900 python -m pytest tests -m gpu -x -q > $OUT/pytest_gpu.log 2>&1
echo "pytest exit $?" >> $OUT/pytest_gpu.log
tail -5 $OUT/pytest_gpu.log

timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $OUT/smoke.log 2>&1
tail -2 $OUT/smoke.log

for wl in cfg3 cfg2 cfg4 cfg5 prod; do
  timeout 600 python bench.py --workload $wl --steps 20 --warmup 5 $( [ $wl != cfg3 ] && echo --no-cpu-baseline ) > $OUT/bench_$wl.json 2> $OUT/bench_$wl.err
  echo "bench $wl exit $?"; tail -c 600 $OUT/bench_$wl.json
done
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > $OUT/bench_reference.json 2> $OUT/bench_reference.err

# launch list of the bench command (cold-cache, serialised: shares only)
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/launches.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --e2e-steps 1 > $OUT/launches_run.log 2>&1

# full capture of the two dominant kernels (after the warm-up launches)
timeout 900 ncu --set full --clock-control none --import-source on -k 'regex:selscan_(fwd_v4|bwd_v2)' -s 6 -c 2 -f -o $OUT/prof_selscan \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline --e2e-steps 1 > $OUT/prof_run.log 2>&1
python tools/pcie_probe.py > $OUT/pcie.txt 2>&1; cat $OUT/pcie.txt
ls -la $OUT
