#!/bin/bash
# Slim N-GPU set: gpu_scale2.sh TAG N "cfg3 cfg4 train tests"
TAG=${1:-scale2}; N=${2:-2}; WHAT=${3:-"cfg3 cfg4"}; OUT=gpurun_out/$TAG; mkdir -p $OUT
run() { timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $1 "${@:2}"; }
for w in $WHAT; do
case $w in
cfg3) run 29511 bench.py --gpus $N --steps 20 --warmup 5 > $OUT/cfg3_batch_n$N.json 2> $OUT/cfg3_batch_n$N.err
  echo "cfg3 batch exit $?"; python -c "import json;d=json.load(open('$OUT/cfg3_batch_n$N.json'));print('cfg3 N=$N', round(d['value']/1e6,2),'Mtok/s', d['ms_per_step'],'ms', d.get('parity_check'), 'e2e', d['e2e']['ms_per_step'])" || tail -5 $OUT/cfg3_batch_n$N.err;;
cfg4) run 29512 bench.py --workload cfg4 --shard channels --gpus $N --steps 20 --warmup 5 --e2e-steps 2 > $OUT/cfg4_channels_n$N.json 2> $OUT/cfg4_channels_n$N.err
  echo "cfg4 channels exit $?"; python -c "import json;d=json.load(open('$OUT/cfg4_channels_n$N.json'));print('cfg4 N=$N', round(d['value']/1e6,2),'Mtok/s', d['ms_per_step'],'ms', d.get('parity_check'), [(k['kernel'],k['avg_ms']) for k in d['kernels'][:6]], d.get('collective'))" || tail -5 $OUT/cfg4_channels_n$N.err;;
train) run 29513 tools/bench_train.py --dtype bf16 > $OUT/cfg5_train_bf16_n$N.json 2> $OUT/cfg5_train_bf16_n$N.err
  echo "cfg5 train exit $?"; python -c "import json;d=json.load(open('$OUT/cfg5_train_bf16_n$N.json'));print({k:d[k] for k in d if 'allreduce' in k or k in ('tokens_per_s','ms_per_step','replicas_identical')})" || tail -5 $OUT/cfg5_train_bf16_n$N.err;;
tests) timeout 600 python -m pytest tests/test_multi_gpu.py -q -m gpu > $OUT/pytest_multi_gpu_n$N.log 2>&1; tail -3 $OUT/pytest_multi_gpu_n$N.log;;
esac
done
