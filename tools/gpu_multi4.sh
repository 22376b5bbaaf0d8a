#!/bin/bash
# N-GPU bench line only.  gpurun --gpus N -- 'bash tools/gpu_multi4.sh tag N'
TAG=${1:-multi}; N=${2:-4}; OUT=gpurun_out/$TAG; mkdir -p $OUT
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29541 \
    bench.py --gpus $N --steps 20 --warmup 5 > $OUT/bench_n$N.json 2> $OUT/bench_n$N.err
python - <<PY
import json
b=json.load(open("$OUT/bench_n$N.json")); print("N=$N", b["value"]/1e6, "Mtok/s", b["ms_per_step"], "ms/step e2e", b["e2e"]["ms_per_step"])
PY
