#!/bin/bash
# first process on a fresh box: bench with settle steps
set -u
OUT=gpurun_out/r02_run37
mkdir -p "$OUT"
export PYTHONUNBUFFERED=1
N=${1:-1}
if [ "$N" = "1" ]; then
  timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > "$OUT/bench_1gpu.json" 2> "$OUT/bench_1gpu.err"
else
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus $N --steps 5 --warmup 3 > "$OUT/bench_${N}gpu.json" 2> "$OUT/bench_${N}gpu.err"
fi
python -c "
import json;d=json.loads(open('$OUT/bench_${N}gpu.json').read().strip().splitlines()[-1]);print($N, d['value'], d['e2e']['value'], d['per_step_ms'], d['config']['settle_ms'])"
