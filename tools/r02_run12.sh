#!/bin/bash
set -u
OUT=gpurun_out/r02_run12
mkdir -p "$OUT"
export PYTHONUNBUFFERED=1
run() { name=$1; shift; echo "== $name: $*"; timeout "$TMO" "$@" > "$OUT/$name" 2>&1; echo "   rc=$? ($(tail -n 1 "$OUT/$name" | cut -c1-300))"; }
TMO=600; run 00_e2e_tests_verbose.txt python -m pytest tests/test_gpu_fitting.py tests/test_gpu_zz_fresh_inputs.py -q -s -k "evaluation_fitting_loss or fitting_loss_fresh or splinenet or open_spline_training"
ls -la "$OUT"
