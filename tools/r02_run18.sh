#!/bin/bash
# f1 (Kronecker fit kernel + optimisers), full GPU suite, bench with per-step times, ncu launch list of the bench command
set -u
OUT=gpurun_out/r02_run18
mkdir -p "$OUT"
export PYTHONUNBUFFERED=1
run() { local name=$1; shift; echo "=== $name: $*" ; timeout ${TMO:-600} "$@" > "$OUT/$name" 2>&1; echo "rc=$? $(tail -n 3 "$OUT/$name" | cut -c1-300)"; }
TMO=600 run 00_f1_tests.txt python -m pytest tests/test_gpu_fitting.py -x -q -m gpu -k "kronecker or cfg1"
TMO=1500 run 10_gpu_suite.txt python -m pytest tests -q -m gpu
TMO=600 run 20_bench.json python bench.py --steps 5 --warmup 3 --no-cpu-baseline
TMO=900 run 30_ncu_launches.txt ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file "$OUT/launches.csv" python bench.py --steps 1 --warmup 1 --no-cpu-baseline
