#!/bin/bash
# Per-kernel roofline benches of the non-scan kernels (warm-up outside the timed window).
TAG=${1:-small1}; OUT=gpurun_out/$TAG; mkdir -p $OUT
(python tools/bench_pscan.py 8 1024 1024 16; python tools/bench_pscan.py 2 1858 1024 16; python tools/bench_pscan.py 32 256 512 16) > $OUT/pscan.txt 2>$OUT/err.txt
(python tools/bench_conv1d.py 16 4096 1536 bf16; python tools/bench_conv1d.py 256 1024 1024 bf16; python tools/bench_conv1d.py 16 4096 1536 f32; python tools/bench_conv1d.py 2 1858 1024 bf16) > $OUT/conv1d.txt 2>>$OUT/err.txt
(python tools/bench_addnorm.py 65536 768 bf16; python tools/bench_addnorm.py 262144 512 bf16; python tools/bench_addnorm.py 262144 512 f32) > $OUT/addnorm.txt 2>>$OUT/err.txt
for f in pscan conv1d addnorm; do python - <<PY
import json
for l in open("$OUT/$f.txt"):
    d=json.loads(l); print({k:v for k,v in d.items() if k not in ("kernels","op","peak_gbs")}, [(k["kernel"],k["avg_ms"],k["frac_of_measured_peak"]) for k in d["kernels"]])
PY
done
