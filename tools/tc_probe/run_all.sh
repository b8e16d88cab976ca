#!/bin/bash
# builds and runs every tcgen05 probe (a few seconds of GPU time); each run is wrapped in `timeout`.
#   gpurun --timeout 300 -- 'bash tools/tc_probe/run_all.sh > gpurun_out/tc_probe.txt 2>&1'
set -u
cd "$(dirname "$0")"
INC="-I ../../parsenet-codebase_b200/csrc"      # the probes include the SHIPPED tc05.cuh (no copy to drift)
nvcc -std=c++17 -gencode arch=compute_100a,code=sm_100a -O2 $INC -o probe probe.cu || exit 1
nvcc -std=c++17 -gencode arch=compute_100a,code=sm_100a -O2 $INC -o probe2 probe2.cu || exit 1
echo "== validated encodings (must be exact): K-major SS / TS"
timeout 20 ./probe 0; timeout 20 ./probe 1
echo "== MN-major B, no swizzle (variant 0: LBO = K direction, SBO = MN direction; variant 1: swapped)"
for m in 2 3; do for v in 0 1; do timeout 20 ./probe $m $v; done; done
echo "== does the tensor core truncate or round full-mantissa operands?  (smem operands, TMEM A operand)"
timeout 20 ./probe 4; timeout 20 ./probe 5
echo "== 128B-swizzled (TMA-style) B tile: K-major read, then the same kind of tile read MN-major"
timeout 20 ./probe 8; timeout 20 ./probe 9
for v in 0 1; do timeout 20 ./probe 10 $v; timeout 20 ./probe 11 $v; done
echo "== CTA pair (tcgen05 cta_group::2): B staged by threads, then by 2-SM TMA loads completing on the leader barrier"
timeout 20 ./probe2 0; timeout 20 ./probe2 1
