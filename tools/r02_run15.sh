#!/bin/bash
# bracketed K-th distance, second version (no atomics in the collect pass, bit-transposed select)
set -u
OUT=gpurun_out/r02_run15
mkdir -p "$OUT"
export PYTHONUNBUFFERED=1
run() { local name=$1; shift; echo "=== $name: $*" ; timeout ${TMO:-600} "$@" > "$OUT/$name" 2>&1; echo "rc=$? $(tail -n 3 "$OUT/$name" | cut -c1-300)"; }
TMO=600 run 00_kth_tests.txt python -m pytest tests/test_gpu_meanshift_tc.py tests/test_gpu_meanshift.py -x -q -m gpu
TMO=300 run 10_kth_ab.txt python tools/exp_ms_kth.py 16 10000 150
TMO=300 run 11_kth_ncu.txt ncu --metrics gpu__time_duration.sum --clock-control none --csv python tools/exp_ms_kth.py 16 10000 150
TMO=600 run 20_bench.json python bench.py --steps 5 --warmup 3 --no-cpu-baseline
