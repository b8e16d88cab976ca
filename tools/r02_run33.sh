#!/bin/bash
# tensor-core filter for C = 256 (64-row CTAs, split parts stacked along the TMEM lanes): A/B, tests, bench
set -u
OUT=gpurun_out/r02_run33
mkdir -p "$OUT"
export PYTHONUNBUFFERED=1
run() { local name=$1; shift; echo "=== $name: $*" ; timeout ${TMO:-600} "$@" > "$OUT/$name" 2>&1; echo "rc=$? $(tail -n 9 "$OUT/$name" | cut -c1-300)"; }
TMO=300 run 00_knn_ab.txt python tools/exp_knn_tc.py 16 feat
TMO=900 run 10_knn_tests.txt python -m pytest tests/test_gpu_knn.py -x -q -m gpu
TMO=600 run 20_bench.json python bench.py --steps 5 --warmup 3 --no-cpu-baseline
