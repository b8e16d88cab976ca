"""round-2 experiment: candidate-buffer capacity of the kNN kernel (PN_KNN_CAP, csrc/knn.cu).  The instruction model in
DESIGN.md section 7 puts ~75 % (C = 6) / ~40 % (C = 64) of the kernel into the quickselect compactions of the per-row
buffers; capacity 256 instead of 128 makes them 3.7x rarer at 1 instead of 2 CTAs per SM, and PN_KNN_SAMPLE=1 starts the
main pass from a sampled admission threshold (~230 instead of ~600 admitted candidates per row).  Indices must be identical."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "parsenet-codebase_b200"))
import torch
from pnb200 import ops

B = int(sys.argv[1]) if len(sys.argv) > 1 else 16
torch.manual_seed(0)
x64 = torch.randn(B, 10000, 64, device="cuda") * 0.3
x6 = torch.randn(B, 10000, 6, device="cuda") * 0.3


def timed(x, metric):
    ops.knn_graph(x, 80, metric); torch.cuda.synchronize()
    best, idx = 1e9, None
    for _ in range(3):
        a = torch.cuda.Event(enable_timing=True); b = torch.cuda.Event(enable_timing=True)
        a.record(); idx = ops.knn_graph(x, 80, metric); b.record(); torch.cuda.synchronize()
        best = min(best, a.elapsed_time(b))
    return best, idx


res = {}
for sample in ("0", "1"):            # PN_KNN_SAMPLE=1: sampled admission threshold (pre-pass over the first 1024 candidates)
    for cap in ("128", "256"):
        os.environ["PN_KNN_CAP"] = cap
        os.environ["PN_KNN_SAMPLE"] = sample
        key = (sample, cap)
        res[key] = (timed(x6, 1), timed(x64, 0))
        same = (torch.equal(res[("0", "128")][0][1], res[key][0][1]), torch.equal(res[("0", "128")][1][1], res[key][1][1]))
        print(f"PN_KNN_SAMPLE={sample} PN_KNN_CAP={cap}: metric-1 C=6 {res[key][0][0]:.2f} ms, C=64 {res[key][1][0]:.2f} ms, "
              f"indices identical to the default: {same}", flush=True)
