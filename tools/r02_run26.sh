#!/bin/bash
# N GPUs: multi-device tests (N = 2) and the driver's scaling command
set -u
N=${1:-2}
OUT=gpurun_out/r02_run26
mkdir -p "$OUT"
export PYTHONUNBUFFERED=1
if [ "$N" = "2" ]; then
  timeout 600 python -m pytest tests/test_gpu_multi_device.py -q -m gpu > "$OUT/00_multi_device_tests.txt" 2>&1; echo "rc=$? $(tail -n 2 "$OUT/00_multi_device_tests.txt")"
fi
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus $N --steps 5 --warmup 3 > "$OUT/bench_${N}gpu.json" 2> "$OUT/bench_${N}gpu.err"
echo "rc=$? $(tail -n 1 "$OUT/bench_${N}gpu.json" | cut -c1-260)"
