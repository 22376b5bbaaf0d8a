#!/bin/bash
# training-step measurements: production shape eager / torch-optimiser / graphed, and the cfg5 stack on N GPUs.  gpurun [--gpus N] -- 'bash tools/gpu_train.sh tag N'
TAG=${1:-train}; N=${2:-1}; OUT=gpurun_out/$TAG; mkdir -p $OUT
if [ "$N" = "1" ]; then
  for v in "--torch-optim" "" "--graph"; do
    timeout 300 python tools/bench_train.py --batch 2 --seq 1858 --layers 6 --dtype f32 --steps 20 --warmup 5 $v > $OUT/prod$(echo $v | tr -d ' -').json 2> $OUT/prod$(echo $v | tr -d ' -').err
    echo "prod $v exit $?"; cat $OUT/prod$(echo $v | tr -d ' -').json | cut -c1-600
  done
  for dt in bf16 tf32; do
    timeout 600 python tools/bench_train.py --dtype $dt > $OUT/cfg5_${dt}_n1.json 2> $OUT/cfg5_${dt}_n1.err; echo "cfg5 $dt exit $?"; cut -c1-900 $OUT/cfg5_${dt}_n1.json; tail -2 $OUT/cfg5_${dt}_n1.err
  done
else
  for dt in bf16; do
    timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29541 tools/bench_train.py --dtype $dt > $OUT/cfg5_${dt}_n$N.json 2> $OUT/cfg5_${dt}_n$N.err
    echo "cfg5 $dt N=$N exit $?"; cut -c1-1200 $OUT/cfg5_${dt}_n$N.json; tail -2 $OUT/cfg5_${dt}_n$N.err
  done
fi
