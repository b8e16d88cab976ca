#!/bin/bash
set -u
OUT=gpurun_out/r02_run9
mkdir -p "$OUT"
export PYTHONUNBUFFERED=1
run() { name=$1; shift; echo "== $name: $*"; timeout "$TMO" "$@" > "$OUT/$name" 2>&1; echo "   rc=$? ($(tail -n 1 "$OUT/$name" | cut -c1-300))"; }
TMO=600; run 00_knn_tests.txt python -m pytest tests/test_gpu_knn.py -q -x
TMO=200; run 05_knn_ab.txt python tools/exp_knn_tma.py 16
TMO=600; run 10_other_tests.txt python -m pytest tests/test_gpu_segnet.py tests/test_gpu_fitting.py tests/test_gpu_fitstage.py -q
TMO=400; run 20_bench.json python bench.py --steps 5 --warmup 3 --no-cpu-baseline
TMO=400; PN_KNN=simt run 21_bench_knn_simt.json python bench.py --steps 5 --warmup 3 --no-cpu-baseline
TMO=400; PROF_STEPS=3 run 30_prof_step.txt python tools/prof_step.py "$OUT/step"
ls -la "$OUT"
