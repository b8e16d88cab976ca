#!/bin/bash
# second GPU call of round 2: new defaults (sparse-row backward, TMA operands, sampled kNN threshold), pinned bench workload
set -u
OUT=gpurun_out/r02_run2
mkdir -p "$OUT"
export PYTHONUNBUFFERED=1
run() { name=$1; shift; echo "== $name: $*"; timeout "$TMO" "$@" > "$OUT/$name" 2>&1; echo "   rc=$? ($(tail -n 1 "$OUT/$name" | cut -c1-200))"; }
TMO=900; run 00_gpu_tests.txt python -m pytest tests -m gpu -q -rxXs
TMO=120; run 01_smoke.txt python -c "import __graft_entry__ as g; g.smoke()"
TMO=400; PROF_STEPS=3 run 10_prof_step.txt python tools/prof_step.py "$OUT/step"
TMO=400; run 20_bench.json python bench.py --steps 5 --warmup 3 --no-cpu-baseline
ls -la "$OUT"
