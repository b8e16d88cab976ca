#!/bin/bash
# host-side trims of the fit stage (vectorised plan, IoU bookkeeping before the last read-back): tests + bench
set -u
OUT=gpurun_out/r02_run32
mkdir -p "$OUT"
export PYTHONUNBUFFERED=1
run() { local name=$1; shift; echo "=== $name: $*" ; timeout ${TMO:-600} "$@" > "$OUT/$name" 2>&1; echo "rc=$? $(tail -n 3 "$OUT/$name" | cut -c1-300)"; }
TMO=900 run 00_tests.txt python -m pytest tests/test_gpu_fitstage.py tests/test_gpu_fitting.py tests/test_gpu_zz_fresh_inputs.py -x -q -m gpu
TMO=600 run 20_bench.json python bench.py --steps 5 --warmup 3 --no-cpu-baseline
