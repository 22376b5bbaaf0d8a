#!/bin/bash
# N-GPU bench only, as the driver launches it (+ the same without NUMA binding).  gpurun --gpus N -- 'bash tools/gpu_multi8.sh tag N'
TAG=${1:-multi}; N=${2:-8}; OUT=gpurun_out/$TAG; mkdir -p $OUT
lscpu | grep -i -E "numa|socket|^cpu\(s\)" > $OUT/lscpu.txt 2>&1; nvidia-smi topo -m > $OUT/topo.txt 2>&1
for mode in bind nobind; do
  extra=""; [ $mode = nobind ] && extra="--no-numa-bind"
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29521 \
      bench.py --gpus $N --steps 20 --warmup 5 $extra > $OUT/bench_n${N}_$mode.json 2> $OUT/bench_n${N}_$mode.err
  echo "$mode exit $?"; python - <<PY
import json
try:
    b=json.load(open("$OUT/bench_n${N}_$mode.json")); print("N=$N $mode", b["value"]/1e6, "Mtok/s", b["ms_per_step"], "ms/step e2e", b["e2e"]["ms_per_step"], "ms cpus", b["e2e"].get("host_cpus"))
except Exception as e:
    print("no json", e); print(open("$OUT/bench_n${N}_$mode.err").read()[-2000:])
PY
done
cat $OUT/lscpu.txt
