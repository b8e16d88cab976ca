#!/bin/bash
set -u
OUT=gpurun_out/r02_run8
mkdir -p "$OUT"
export PYTHONUNBUFFERED=1
run() { name=$1; shift; echo "== $name: $*"; timeout "$TMO" "$@" > "$OUT/$name" 2>&1; echo "   rc=$? ($(tail -n 1 "$OUT/$name" | cut -c1-300))"; }
TMO=300; run 00_linear_tests.txt python -m pytest tests/test_gpu_linear_tc.py -q -s
TMO=600; run 01_segnet_tests.txt python -m pytest tests/test_gpu_segnet.py tests/test_gpu_fitting.py tests/test_gpu_zz_fresh_inputs.py -q
TMO=400; run 20_bench.json python bench.py --steps 5 --warmup 3 --no-cpu-baseline
TMO=400; PN_LINEAR_BWD=simt run 21_bench_bwd_simt.json python bench.py --steps 5 --warmup 3 --no-cpu-baseline
TMO=400; PROF_STEPS=3 run 30_prof_step.txt python tools/prof_step.py "$OUT/step"
ls -la "$OUT"
