#!/bin/bash
# First GPU call of round 2: everything that was written after the GPU budget of round 1 ran out, in order of risk, each
# step under its own timeout, all output under gpurun_out/r02_first/.
#   gpurun --timeout 2400 -- 'bash tools/r02_runbook.sh'      (typically ~20 min: tests 5, six bench runs 9, the rest 5)
# The big one is 24_bench_sparse_bwd.json: the mean-shift backward (55 % of the device time of a step) only has to run for
# the <= 49 centre rows the loss actually sees (exact).
# Reading order afterwards: 00_gpu_tests.txt (XPASS = promote the test, XFAIL = read the assertion), 10_tma_cg1.txt,
# 11_tma_cg12.txt (bit-identical? faster?), 20_bench_default.json vs 21_bench_tma*.json.
set -u
OUT=gpurun_out/r02_first
mkdir -p "$OUT"
export PYTHONUNBUFFERED=1
run() { name=$1; shift; echo "== $name: $*"; timeout "$TMO" "$@" > "$OUT/$name" 2>&1; echo "   rc=$? ($(tail -n 1 "$OUT/$name" | cut -c1-160))"; }

TMO=900; run 00_gpu_tests.txt python -m pytest tests -m gpu -q -rxXs -x
TMO=120; run 01_smoke.txt python -c "import __graft_entry__ as g; g.smoke()"
TMO=200; run 02_tc_probes.txt bash tools/tc_probe/run_all.sh
# experimental kernels (they trap after ~2 s instead of hanging): 1-CTA TMA first, small then full size, then CTA pairs
TMO=120; PN_EXP_CGS=1 run 10_tma_cg1_small.txt python tools/exp_ms_tma.py 2 1000
TMO=180; PN_EXP_CGS=1 run 10_tma_cg1.txt python tools/exp_ms_tma.py 16 10000
TMO=120; PN_EXP_CGS=2 run 11_tma_cg2_small.txt python tools/exp_ms_tma.py 2 1000
TMO=180; PN_EXP_CGS=1,2 run 11_tma_cg12.txt python tools/exp_ms_tma.py 16 10000
TMO=300; PN_RUN_EXPERIMENTAL=1 run 12_tma_tests.txt python -m pytest tests/test_gpu_meanshift_tc.py -q -k tma
TMO=300; PN_RUN_EXPERIMENTAL=1 run 13_fit_batched_test.txt python -m pytest tests/test_gpu_zz_first_run.py -q -k batched
TMO=300; PN_RUN_EXPERIMENTAL=1 run 14_sparse_bwd_test.txt python -m pytest tests/test_gpu_zz_first_run.py -q -k sparse_row
TMO=300; PN_MS_SPARSE_BWD=1 run 15_e2e_tests_sparse_bwd.txt python -m pytest tests/test_gpu_fitting.py -q -k evaluation_fitting_loss
# the bench with and without them (same box, back to back)
TMO=400; run 20_bench_default.json python bench.py --steps 5 --warmup 3 --no-cpu-baseline
TMO=400; PN_MS_TMA=1 run 21_bench_tma_cg1.json python bench.py --steps 5 --warmup 3 --no-cpu-baseline
TMO=400; PN_MS_TMA=1 PN_MS_TMA_CG=2 run 22_bench_tma_cg2.json python bench.py --steps 5 --warmup 3 --no-cpu-baseline
TMO=400; PN_MS_SPARSE_BWD=1 run 24_bench_sparse_bwd.json python bench.py --steps 5 --warmup 3 --no-cpu-baseline
TMO=400; PN_MS_SPARSE_BWD=1 PN_FIT_BATCHED=1 run 25_bench_sparse_bwd_fit_batched.json python bench.py --steps 5 --warmup 3 --no-cpu-baseline
TMO=400; PN_FIT_BATCHED=1 run 23_bench_fit_batched.json python bench.py --steps 5 --warmup 3 --no-cpu-baseline
TMO=120; run 30_ms_bwd_sweep.txt python tools/exp_ms_bwd.py 0,128,160
TMO=120; run 31_knn_cap.txt python tools/exp_knn_cap.py 16
ls -la "$OUT"
