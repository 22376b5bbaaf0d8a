#!/bin/bash
# BASELINE config 4 as specified: ONE sequence, ED-channel-sharded over N GPUs.  gpurun --gpus N -- 'bash tools/gpu_cfg4.sh tag N'
TAG=${1:-cfg4}; N=${2:-2}; OUT=gpurun_out/$TAG; mkdir -p $OUT
timeout 300 python bench.py --workload cfg4 --steps 10 --warmup 3 --no-cpu-baseline --e2e-steps 1 > $OUT/n1.json 2> $OUT/n1.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29531 \
    bench.py --workload cfg4 --shard channels --gpus $N --steps 10 --warmup 3 --e2e-steps 1 > $OUT/n$N.json 2> $OUT/n$N.err
python - <<PY
import json
for n in (1, $N):
    try:
        d=json.load(open("$OUT/n%d.json" % n)); print("N=%d" % n, round(d["value"]/1e6,2), "Mtok/s", d["ms_per_step"], "ms/step", d["scaling"], d["config"]["sharding"])
    except Exception as e:
        print("N=%d no json" % n, e); print(open("$OUT/n%d.err" % n).read()[-1500:])
PY
