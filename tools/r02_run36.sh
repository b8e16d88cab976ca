#!/bin/bash
# bench with the in-process NVML clock sampler: is the first timed loop flat now?  (two runs)
set -u
OUT=gpurun_out/r02_run36
mkdir -p "$OUT"
export PYTHONUNBUFFERED=1
for r in 1 2; do
  timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > "$OUT/bench_$r.json" 2> "$OUT/bench_$r.err"
  python -c "
import json;d=json.loads(open('$OUT/bench_$r.json').read().strip().splitlines()[-1]);print($r, d['value'], d['e2e']['value'], d['per_step_ms'], d['clocks'])"
done
timeout 600 python -m pytest tests/test_gpu_meanshift_tc.py -x -q -m gpu -k argsel 2>&1 | tail -2
