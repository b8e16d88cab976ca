#!/bin/bash
# per-kernel times of the bracketed kNN paths
set -u
OUT=gpurun_out/r02_run23
mkdir -p "$OUT"
export PYTHONUNBUFFERED=1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file "$OUT/knn_launches.csv" python tools/exp_knn_tc.py 16 feat > "$OUT/00.txt" 2>&1
echo "rc=$?"; tail -3 "$OUT/00.txt"
