#!/bin/bash
# third GPU call of round 2: batched fit stage
set -u
OUT=gpurun_out/r02_run3
mkdir -p "$OUT"
export PYTHONUNBUFFERED=1
run() { name=$1; shift; echo "== $name: $*"; timeout "$TMO" "$@" > "$OUT/$name" 2>&1; echo "   rc=$? ($(tail -n 1 "$OUT/$name" | cut -c1-200))"; }
TMO=600; run 00_fitstage_tests.txt python -m pytest tests/test_gpu_fitstage.py tests/test_gpu_fitting.py tests/test_gpu_zz_fresh_inputs.py -q -x -k "fitstage or fit_stage or fit_solve or batched or evaluation_fitting_loss or fitting_loss_fresh"
TMO=900; run 01_gpu_tests.txt python -m pytest tests -m gpu -q -rxXs
TMO=400; PROF_STEPS=3 run 10_prof_step.txt python tools/prof_step.py "$OUT/step"
TMO=400; run 20_bench.json python bench.py --steps 5 --warmup 3 --no-cpu-baseline
TMO=400; PN_FIT_STAGE=loop run 21_bench_loop.json python bench.py --steps 5 --warmup 3 --no-cpu-baseline
ls -la "$OUT"
