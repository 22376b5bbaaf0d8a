#!/bin/bash
TAG=${1:-r2g}; OUT=gpurun_out/$TAG; mkdir -p $OUT
timeout 900 python -m pytest tests -m gpu -q > $OUT/pytest_gpu.log 2>&1; echo "pytest exit $?" | tee -a $OUT/pytest_gpu.log
tail -8 $OUT/pytest_gpu.log
python tools/bench_pscan.py 8 1024 1024 16 > $OUT/pscan.txt 2>&1; python tools/bench_pscan.py 2 1858 1024 16 >> $OUT/pscan.txt 2>&1; python tools/bench_pscan.py 32 256 512 16 >> $OUT/pscan.txt 2>&1; cat $OUT/pscan.txt
timeout 300 python tools/bench_train.py --batch 2 --seq 1858 --layers 6 --dtype tf32 --steps 20 --warmup 5 --graph > $OUT/prod_graph_tf32.json 2> $OUT/prod_graph_tf32.err; cut -c1-400 $OUT/prod_graph_tf32.json
timeout 300 python tools/bench_train.py --batch 2 --seq 1858 --layers 6 --dtype bf16 --steps 20 --warmup 5 --graph > $OUT/prod_graph_bf16.json 2> $OUT/prod_graph_bf16.err; cut -c1-400 $OUT/prod_graph_bf16.json
