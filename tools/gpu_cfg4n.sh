#!/bin/bash
# BASELINE config 4 as specified (ONE sequence, ED-channel-sharded over N GPUs), one line per N.  gpurun --gpus N -- 'bash tools/gpu_cfg4n.sh tag N'
TAG=${1:-cfg4n}; N=${2:-2}; OUT=gpurun_out/$TAG; mkdir -p $OUT
if [ "$N" = "1" ]; then
  timeout 300 python bench.py --workload cfg4 --steps 20 --warmup 5 --no-cpu-baseline --e2e-steps 2 > $OUT/cfg4_channels_n1.json 2> $OUT/cfg4_channels_n1.err
else
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29531 \
    bench.py --workload cfg4 --shard channels --gpus $N --steps 20 --warmup 5 --e2e-steps 2 > $OUT/cfg4_channels_n$N.json 2> $OUT/cfg4_channels_n$N.err
fi
python -c "import json;d=json.load(open('$OUT/cfg4_channels_n$N.json'));print('cfg4 N=$N', round(d['value']/1e6,2),'Mtok/s', d['ms_per_step'],'ms', d.get('parity_check'), [(k['kernel'],k['avg_ms']) for k in d['kernels'][:6]])"
