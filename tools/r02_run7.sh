#!/bin/bash
set -u
OUT=gpurun_out/r02_run7
mkdir -p "$OUT"
export PYTHONUNBUFFERED=1
run() { name=$1; shift; echo "== $name: $*"; timeout "$TMO" "$@" > "$OUT/$name" 2>&1; echo "   rc=$? ($(tail -n 1 "$OUT/$name" | cut -c1-300))"; }
TMO=300; run 00_diag_spline.txt python tests/diag_fitstage_spline.py
TMO=600; run 01_tests.txt python -m pytest tests/test_gpu_fitstage.py tests/test_gpu_fitting.py tests/test_gpu_zz_fresh_inputs.py -q
ls -la "$OUT"
