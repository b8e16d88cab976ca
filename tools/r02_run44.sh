#!/bin/bash
# last check of the final tree: GPU suite + bench line
set -u
OUT=gpurun_out/r02_run44
mkdir -p "$OUT"
export PYTHONUNBUFFERED=1
timeout 900 python -m pytest tests -m gpu -q > "$OUT/00_gpu_tests.txt" 2>&1; echo "rc=$? $(tail -n 1 "$OUT/00_gpu_tests.txt")"
timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > "$OUT/20_bench.json" 2> "$OUT/20_bench.err"; echo "rc=$? $(tail -n 1 "$OUT/20_bench.json" | cut -c1-240)"
