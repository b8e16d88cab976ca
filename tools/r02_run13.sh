#!/bin/bash
# 8 GPUs: the driver's scaling command for N = 8 (weak scaling headline + strong figure + per-rank times)
set -u
OUT=gpurun_out/r02_run13
mkdir -p "$OUT"
export PYTHONUNBUFFERED=1
N=${1:-8}
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus $N --steps 5 --warmup 3 > "$OUT/bench_${N}gpu.json" 2> "$OUT/bench_${N}gpu.err"
echo "rc=$? $(tail -n 1 "$OUT/bench_${N}gpu.json" | cut -c1-200)"
