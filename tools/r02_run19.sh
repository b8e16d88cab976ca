#!/bin/bash
# device matching (IoU cost + Hungarian), f1 tests again, bench (inference path with the device matching)
set -u
OUT=gpurun_out/r02_run19
mkdir -p "$OUT"
export PYTHONUNBUFFERED=1
run() { local name=$1; shift; echo "=== $name: $*" ; timeout ${TMO:-600} "$@" > "$OUT/$name" 2>&1; echo "rc=$? $(tail -n 3 "$OUT/$name" | cut -c1-300)"; }
TMO=600 run 00_tests.txt python -m pytest tests/test_gpu_assign.py tests/test_gpu_fitting.py -x -q -m gpu -k "assign or iou or hungarian or match or kronecker"
TMO=600 run 20_bench.json python bench.py --steps 5 --warmup 3 --no-cpu-baseline
