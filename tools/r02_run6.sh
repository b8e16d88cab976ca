#!/bin/bash
set -u
OUT=gpurun_out/r02_run6
mkdir -p "$OUT"
export PYTHONUNBUFFERED=1
run() { name=$1; shift; echo "== $name: $*"; timeout "$TMO" "$@" > "$OUT/$name" 2>&1; echo "   rc=$? ($(tail -n 1 "$OUT/$name" | cut -c1-300))"; }
TMO=600; run 00_new_tests.txt python -m pytest tests/test_gpu_fitstage.py tests/test_gpu_fitting.py -q
TMO=400; run 05_parity_bounds.txt python tests/diag_parity_bounds.py
TMO=900; run 20_bench_full.json python bench.py --steps 5 --warmup 3
ls -la "$OUT"
