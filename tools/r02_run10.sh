#!/bin/bash
set -u
OUT=gpurun_out/r02_run10
mkdir -p "$OUT"
export PYTHONUNBUFFERED=1
run() { name=$1; shift; echo "== $name: $*"; timeout "$TMO" "$@" > "$OUT/$name" 2>&1; echo "   rc=$? ($(tail -n 1 "$OUT/$name" | cut -c1-300))"; }
TMO=600; run 00_knn_tests.txt python -m pytest tests/test_gpu_knn.py -q -x
TMO=200; run 05_knn_ab.txt python tools/exp_knn_tma.py 16
TMO=400; run 20_bench.json python bench.py --steps 5 --warmup 3 --no-cpu-baseline
TMO=300; run 40_ncu_knn.txt ncu --set full --clock-control none --import-source on -k regex:knn_tma -c 3 -o "$OUT/knn_tma" -f python tools/exp_knn_tma.py 16
ls -la "$OUT"
