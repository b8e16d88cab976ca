#!/bin/bash
# full GPU suite, smoke, full bench (CPU arm + GPU-eager baseline + inference path), reference arm
set -u
OUT=gpurun_out/r02_run42
mkdir -p "$OUT"
export PYTHONUNBUFFERED=1
run() { local name=$1; shift; echo "=== $name: $*" ; timeout ${TMO:-600} "$@" > "$OUT/$name" 2>&1; echo "rc=$? $(tail -n 3 "$OUT/$name" | cut -c1-300)"; }
TMO=1500 run 00_gpu_tests.txt python -m pytest tests -m gpu -q -rxXs --durations=5
TMO=200 run 01_smoke.txt python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')"
TMO=900 run 20_bench_full.json python bench.py --steps 5 --warmup 3
TMO=600 run 21_bench_reference.json python bench.py --impl reference --steps 1 --warmup 0
