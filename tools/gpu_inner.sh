#!/bin/bash
TAG=${1:-inner1}; OUT=gpurun_out/$TAG; mkdir -p $OUT
timeout 600 python -m pytest tests -m gpu -x -q -k "block or mamba or inner or graphed or train or optim or decoder or layernorm" > $OUT/quick.log 2>&1; tail -5 $OUT/quick.log
timeout 600 python tools/bench_train.py --dtype bf16 > $OUT/train_cfg5_bf16.json 2> $OUT/train.err; python -c "import json;d=json.load(open('$OUT/train_cfg5_bf16.json'));print('cfg5 train', d['ms_per_step'], d['tokens_per_s'], d['library_kernels_ms'])" || tail -5 $OUT/train.err
timeout 300 python tools/bench_train.py --batch 2 --seq 1858 --layers 6 --dtype bf16 --steps 20 --warmup 5 --graph > $OUT/train_prod_bf16.json 2>> $OUT/train.err; python -c "import json;d=json.load(open('$OUT/train_prod_bf16.json'));print('prod graph bf16', d['ms_per_step'])" || tail -5 $OUT/train.err
timeout 300 python tools/bench_train.py --batch 2 --seq 1858 --layers 6 --dtype f32 --steps 20 --warmup 5 > $OUT/train_prod_f32.json 2>> $OUT/train.err; python -c "import json;d=json.load(open('$OUT/train_prod_f32.json'));print('prod eager f32', d['ms_per_step'], d['library_kernels_ms_per_step'])" || tail -5 $OUT/train.err
