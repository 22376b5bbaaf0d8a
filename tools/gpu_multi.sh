#!/bin/bash
# N-GPU bench exactly as the driver launches it.  gpurun --gpus N -- 'bash tools/gpu_multi.sh tag N'
TAG=${1:-multi}; N=${2:-2}; OUT=gpurun_out/$TAG; mkdir -p $OUT
nvidia-smi --query-gpu=index,name --format=csv > $OUT/gpus.txt
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 \
    bench.py --gpus $N --steps 20 --warmup 5 > $OUT/bench_n$N.json 2> $OUT/bench_n$N.err
echo "exit $?"; tail -c 1500 $OUT/bench_n$N.json; tail -3 $OUT/bench_n$N.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 \
    bench.py --impl reference --gpus $N --steps 2 --warmup 1 > $OUT/ref_n$N.json 2> $OUT/ref_n$N.err
echo "ref exit $?"; tail -c 400 $OUT/ref_n$N.json
timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > $OUT/bench_n1.json 2> $OUT/bench_n1.err
python - <<PY
import json
a=json.load(open("$OUT/bench_n1.json")); b=json.load(open("$OUT/bench_n$N.json"))
print("N=1", a["value"]/1e6, "Mtok/s e2e", a["e2e"]["ms_per_step"], "| N=$N", b["value"]/1e6, "Mtok/s e2e", b["e2e"]["ms_per_step"], "eff", b["value"]/a["value"]/$N)
PY
