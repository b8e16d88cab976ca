#!/bin/bash
# bracketed K-th distance with 2048-entry lists (the bench's K = 250), ncu of the collect pass, bench
set -u
OUT=gpurun_out/r02_run17
mkdir -p "$OUT"
export PYTHONUNBUFFERED=1
run() { local name=$1; shift; echo "=== $name: $*" ; timeout ${TMO:-600} "$@" > "$OUT/$name" 2>&1; echo "rc=$? $(tail -n 3 "$OUT/$name" | cut -c1-300)"; }
TMO=600 run 00_kth_tests.txt python -m pytest tests/test_gpu_meanshift_tc.py tests/test_gpu_meanshift.py -x -q -m gpu
TMO=300 run 10_kth_ab.txt python tools/exp_ms_kth.py 16 10000 250
TMO=300 run 11_kth_ncu.txt ncu --metrics gpu__time_duration.sum --clock-control none --csv python tools/exp_ms_kth.py 16 10000 250
TMO=400 run 12_kth_ncu_full.txt ncu --set full --clock-control none --import-source on -k regex:ms_kth_collect -c 2 -o "$OUT/kth_collect" -f python tools/exp_ms_kth.py 16 10000 250
TMO=600 run 20_bench.json python bench.py --steps 5 --warmup 3 --no-cpu-baseline
