#!/bin/bash
# fifth GPU call of round 2: new kernels (weights, grid losses), parity diagnostics, full GPU suite, bench with GPU-eager baseline
set -u
OUT=gpurun_out/r02_run5
mkdir -p "$OUT"
export PYTHONUNBUFFERED=1
run() { name=$1; shift; echo "== $name: $*"; timeout "$TMO" "$@" > "$OUT/$name" 2>&1; echo "   rc=$? ($(tail -n 1 "$OUT/$name" | cut -c1-300))"; }
TMO=600; run 00_new_tests.txt python -m pytest tests/test_gpu_fitstage.py -q
TMO=300; run 05_parity_bounds.txt python tests/diag_parity_bounds.py
TMO=1200; run 10_gpu_tests.txt python -m pytest tests -m gpu -q -rxXs --durations=8
TMO=900; run 20_bench_full.json python bench.py --steps 5 --warmup 3
TMO=400; PROF_STEPS=3 run 30_prof_step.txt python tools/prof_step.py "$OUT/step"
ls -la "$OUT"
