#!/bin/bash
# full GPU suite + bench + step profile with the round-2 kernels (TMA kNN, tcgen05 MLP backward, batched fit stage)
set -u
OUT=gpurun_out/r02_run11
mkdir -p "$OUT"
export PYTHONUNBUFFERED=1
run() { name=$1; shift; echo "== $name: $*"; timeout "$TMO" "$@" > "$OUT/$name" 2>&1; echo "   rc=$? ($(tail -n 1 "$OUT/$name" | cut -c1-300))"; }
TMO=1200; run 00_gpu_tests.txt python -m pytest tests -m gpu -q -rxXs --durations=5
TMO=120; run 01_smoke.txt python -c "import __graft_entry__ as g; g.smoke()"
TMO=200; run 05_knn_ab.txt python tools/exp_knn_tma.py 16
TMO=900; run 20_bench_full.json python bench.py --steps 5 --warmup 3
TMO=400; PROF_STEPS=3 run 30_prof_step.txt python tools/prof_step.py "$OUT/step"
ls -la "$OUT"
