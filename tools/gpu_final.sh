#!/bin/bash
# Round artifacts in ONE gpurun call: parity tests, smoke, every bench line + reference arm, ncu launch list, full captures of the
# two scan kernels (-> traffic), per-kernel roofline benches, training-step lines.   gpurun --timeout 2400 -- 'bash tools/gpu_final.sh tag'
TAG=${1:-r02b}; OUT=gpurun_out/$TAG; mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $OUT/gpu.txt 2>&1
timeout 900 python -m pytest tests -m gpu -q > $OUT/pytest_gpu.log 2>&1; echo "pytest exit $?" >> $OUT/pytest_gpu.log; tail -3 $OUT/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $OUT/smoke.log 2>&1; tail -2 $OUT/smoke.log
for wl in cfg3 cfg1 cfg2 cfg4 cfg5 prod; do
  timeout 600 python bench.py --workload $wl --steps 20 --warmup 5 $( [ $wl != cfg3 ] && echo --no-cpu-baseline ) > $OUT/bench_$wl.json 2> $OUT/bench_$wl.err
  echo "bench $wl exit $?"; python -c "import json;d=json.load(open('$OUT/bench_$wl.json'));print('$wl', d['ms_per_step'], round(d['value']/1e6,2), d['roofline']['frac'], d['roofline']['step_frac'], [(k['kernel'],k['avg_ms']) for k in d['kernels'][:4]], 'e2e', d['e2e']['ms_per_step'])"
done
timeout 600 python bench.py --impl reference --steps 3 --warmup 3 > $OUT/bench_reference.json 2> $OUT/bench_reference.err; cut -c1-300 $OUT/bench_reference.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/launches.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --e2e-steps 1 > $OUT/launches_run.log 2>&1
for wl in cfg3 cfg5; do
timeout 900 ncu --set full --clock-control none --import-source on -k 'regex:selscan_(fwd_v4|bwd_chain)' -s 6 -c 2 -f -o $OUT/prof_$wl \
    python bench.py --workload $wl --steps 1 --warmup 3 --no-cpu-baseline --e2e-steps 1 > $OUT/prof_run_$wl.log 2>&1
done
python tools/bench_pscan.py 8 1024 1024 16 > $OUT/pscan.txt 2>&1; python tools/bench_pscan.py 2 1858 1024 16 >> $OUT/pscan.txt 2>&1
python tools/bench_conv1d.py > $OUT/conv1d.txt 2>&1; python tools/bench_addnorm.py > $OUT/addnorm.txt 2>&1; python tools/bench_decode.py > $OUT/decode.txt 2>&1
for v in "--torch-optim" "" "--graph"; do
  timeout 300 python tools/bench_train.py --batch 2 --seq 1858 --layers 6 --dtype f32 --steps 20 --warmup 5 $v >> $OUT/train_prod.txt 2>> $OUT/train.err
done
timeout 300 python tools/bench_train.py --batch 2 --seq 1858 --layers 6 --dtype tf32 --steps 20 --warmup 5 --graph >> $OUT/train_prod.txt 2>> $OUT/train.err
timeout 300 python tools/bench_train.py --batch 2 --seq 1858 --layers 6 --dtype bf16 --steps 20 --warmup 5 --graph >> $OUT/train_prod.txt 2>> $OUT/train.err
timeout 600 python tools/bench_train.py --dtype bf16 > $OUT/train_cfg5_bf16_n1.json 2>> $OUT/train.err
python tools/pcie_probe.py > $OUT/pcie.txt 2>&1; cat $OUT/pcie.txt
ls $OUT
