#!/bin/bash
# 2-GPU correctness: NCCL parity tests + bench.py's own sharded parity check in both sharding modes.
TAG=${1:-r2f}; N=${2:-2}; OUT=gpurun_out/$TAG; mkdir -p $OUT
nvidia-smi --query-gpu=index,name --format=csv > $OUT/gpus.txt
timeout 600 python -m pytest tests/test_multi_gpu.py -m gpu -q -rA > $OUT/pytest_multi_gpu.log 2>&1; echo "pytest exit $?" | tee -a $OUT/pytest_multi_gpu.log
tail -8 $OUT/pytest_multi_gpu.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 \
    bench.py --gpus $N --steps 10 --warmup 3 > $OUT/bench_batch_n$N.json 2> $OUT/bench_batch_n$N.err
echo "batch exit $?"; python -c "import json;d=json.load(open('$OUT/bench_batch_n$N.json'));print(d['value']/1e6, d['ms_per_step'], d.get('parity_check'), d['e2e']['ms_per_step'])"; tail -3 $OUT/bench_batch_n$N.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 \
    bench.py --workload cfg4 --shard channels --gpus $N --steps 10 --warmup 3 --e2e-steps 1 > $OUT/bench_chan_n$N.json 2> $OUT/bench_chan_n$N.err
echo "channels exit $?"; python -c "import json;d=json.load(open('$OUT/bench_chan_n$N.json'));print(d['value']/1e6, d['ms_per_step'], d.get('parity_check'))"; tail -3 $OUT/bench_chan_n$N.err
