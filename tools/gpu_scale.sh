#!/bin/bash
# N-GPU measurement set: cfg3 batch-sharded (driver launch line), cfg4 channel-sharded, cfg5 stack training step.  gpurun --gpus N -- 'bash tools/gpu_scale.sh tag N'
TAG=${1:-scale}; N=${2:-2}; OUT=gpurun_out/$TAG; mkdir -p $OUT
run() { timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $1 "${@:2}"; }
run 29511 bench.py --gpus $N --steps 20 --warmup 5 > $OUT/cfg3_batch_n$N.json 2> $OUT/cfg3_batch_n$N.err
echo "cfg3 batch exit $?"; python -c "import json;d=json.load(open('$OUT/cfg3_batch_n$N.json'));print('cfg3 N=$N', round(d['value']/1e6,2),'Mtok/s', d['ms_per_step'],'ms', d.get('parity_check'), 'e2e', d['e2e']['ms_per_step'])"
run 29512 bench.py --workload cfg4 --shard channels --gpus $N --steps 20 --warmup 5 --e2e-steps 2 > $OUT/cfg4_channels_n$N.json 2> $OUT/cfg4_channels_n$N.err
echo "cfg4 channels exit $?"; python -c "import json;d=json.load(open('$OUT/cfg4_channels_n$N.json'));print('cfg4 N=$N', round(d['value']/1e6,2),'Mtok/s', d['ms_per_step'],'ms', d.get('parity_check'), [(k['kernel'],k['avg_ms']) for k in d['kernels'][:6]])"
run 29513 tools/bench_train.py --dtype bf16 > $OUT/cfg5_train_bf16_n$N.json 2> $OUT/cfg5_train_bf16_n$N.err
echo "cfg5 train exit $?"; cut -c1-250 $OUT/cfg5_train_bf16_n$N.json; python -c "import json;d=json.load(open('$OUT/cfg5_train_bf16_n$N.json'));print({k:d[k] for k in d if 'allreduce' in k or k in ('tokens_per_s','ms_per_step','replicas_identical')})"
run 29514 tools/pcie_probe.py > $OUT/pcie_n$N.txt 2> $OUT/pcie_n$N.err; cat $OUT/pcie_n$N.txt
