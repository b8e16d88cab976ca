#!/bin/bash
# 2 GPUs: why does the resident loop slow down step by step?  (a) 8 steps, (b) without the nvidia-smi sampler
set -u
OUT=gpurun_out/r02_run27
mkdir -p "$OUT"
export PYTHONUNBUFFERED=1
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 2 --steps 8 --warmup 3 > "$OUT/a.json" 2> "$OUT/a.err"
python -c "
import json;d=json.loads(open('$OUT/a.json').read().strip().splitlines()[-1]);print('a', d['value'], d['e2e']['value'], d['per_step_ms'])"
PN_BENCH_NO_SAMPLER=1 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29514 bench.py --gpus 2 --steps 8 --warmup 3 > "$OUT/b.json" 2> "$OUT/b.err"
python -c "
import json;d=json.loads(open('$OUT/b.json').read().strip().splitlines()[-1]);print('b', d['value'], d['e2e']['value'], d['per_step_ms'])"
nproc; free -g | head -2
