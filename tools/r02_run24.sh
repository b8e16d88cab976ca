#!/bin/bash
# knn_tc final kernel with conflict-only exact costs: tests, A/B, step profile, bench
set -u
OUT=gpurun_out/r02_run24
mkdir -p "$OUT"
export PYTHONUNBUFFERED=1
run() { local name=$1; shift; echo "=== $name: $*" ; timeout ${TMO:-600} "$@" > "$OUT/$name" 2>&1; echo "rc=$? $(tail -n 8 "$OUT/$name" | cut -c1-300)"; }
TMO=900 run 10_knn_tests.txt python -m pytest tests/test_gpu_knn.py tests/test_gpu_segnet.py -x -q -m gpu
TMO=300 run 00_knn_ab.txt python tools/exp_knn_tc.py 16 feat
TMO=600 run 20_bench.json python bench.py --steps 5 --warmup 3 --no-cpu-baseline
TMO=400 PROF_STEPS=3 run 30_prof_step.txt python tools/prof_step.py "$OUT/step"
