#!/bin/bash
# fourth GPU call of round 2: bench line with CPU arm (subprocess) + GPU-eager baseline, ncu of the dominant kernel, launch list
set -u
OUT=gpurun_out/r02_run4
mkdir -p "$OUT"
export PYTHONUNBUFFERED=1
run() { name=$1; shift; echo "== $name: $*"; timeout "$TMO" "$@" > "$OUT/$name" 2>&1; echo "   rc=$? ($(tail -n 1 "$OUT/$name" | cut -c1-300))"; }
TMO=300; run 00_fitstage_tests.txt python -m pytest tests/test_gpu_fitstage.py -q
TMO=900; run 10_bench_full.json python bench.py --steps 5 --warmup 3
TMO=300; run 20_ncu_ms_fwd.txt ncu --set full --clock-control none --import-source on -k regex:ms_fwd_tma -s 2 -c 2 -o "$OUT/ms_fwd_tma" -f python tools/prof_ms.py 16 2
TMO=600; run 30_ncu_launches.txt ncu --metrics gpu__time_duration.sum --clock-control none -k regex:pn:: --csv --log-file "$OUT/launches.csv" python bench.py --steps 1 --warmup 1 --no-cpu-baseline
ls -la "$OUT"
