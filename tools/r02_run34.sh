#!/bin/bash
# TMA-fed nms arg-selects: tests, bench, step profile
set -u
OUT=gpurun_out/r02_run34
mkdir -p "$OUT"
export PYTHONUNBUFFERED=1
run() { local name=$1; shift; echo "=== $name: $*" ; timeout ${TMO:-600} "$@" > "$OUT/$name" 2>&1; echo "rc=$? $(tail -n 4 "$OUT/$name" | cut -c1-300)"; }
TMO=900 run 00_tests.txt python -m pytest tests/test_gpu_meanshift_tc.py tests/test_gpu_meanshift.py tests/test_gpu_fitting.py -x -q -m gpu
TMO=600 run 20_bench.json python bench.py --steps 5 --warmup 3 --no-cpu-baseline
TMO=400 PROF_STEPS=3 run 30_prof_step.txt python tools/prof_step.py "$OUT/step"
