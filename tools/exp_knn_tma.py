"""A/B of the two kNN kernels (csrc/knn.cu loader-thread kernel vs csrc/knn_tma.cu TMA-staged kernel) at the shapes of the
bench step: seg-net layers (N = 10^4, k = 80, C = 64) and SplineNet layers (N = 5000, k = 10, C = 64 / 128 / 256), B = 16."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "parsenet-codebase_b200"))
import torch
from pnb200.cabi import call

B = int(sys.argv[1]) if len(sys.argv) > 1 else 16
torch.manual_seed(0)
st = torch.cuda.current_stream().cuda_stream
print("| N | C | k | loader-thread kernel ms | TMA kernel ms | speed-up | identical |\n|---|---|---|---:|---:|---:|---|")
for N, C, k in ((10000, 64, 80), (5000, 64, 10), (5000, 128, 10), (5000, 256, 10)):
    x = torch.randn(B, N, C, device="cuda") * 0.3
    res = {}
    for name in ("pn_knn", "pn_knn_tma"):
        idx = torch.empty(B, N, k, dtype=torch.int32, device="cuda")
        ws = torch.empty(B * N, device="cuda")
        best = 1e9
        for it in range(4):
            a = torch.cuda.Event(enable_timing=True); b = torch.cuda.Event(enable_timing=True)
            a.record()
            call(name, x.data_ptr(), B, N, C, C, k, 0, idx.data_ptr(), 0, None, ws.data_ptr(), st)
            b.record(); torch.cuda.synchronize()
            if it:
                best = min(best, a.elapsed_time(b))
        res[name] = (best, idx)
    same = torch.equal(res["pn_knn"][1], res["pn_knn_tma"][1])
    print(f"| {N} | {C} | {k} | {res['pn_knn'][0]:.2f} | {res['pn_knn_tma'][0]:.2f} | {res['pn_knn'][0] / res['pn_knn_tma'][0]:.2f}x | {same} |",
          flush=True)
