#!/bin/bash
# 2 GPUs: tensors on a device other than the current one; 2-rank bench (weak + strong figure, per-rank times)
set -u
OUT=gpurun_out/r02_run6b
mkdir -p "$OUT"
export PYTHONUNBUFFERED=1
run() { name=$1; shift; echo "== $name: $*"; timeout "$TMO" "$@" > "$OUT/$name" 2>&1; echo "   rc=$? ($(tail -n 1 "$OUT/$name" | cut -c1-300))"; }
TMO=300; run 00_multi_device_tests.txt python -m pytest tests/test_gpu_multi_device.py -q -rs
TMO=600; run 10_bench_2gpu.json python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 5 --warmup 3
ls -la "$OUT"
