"""A/B of the TMA-staged exact kNN kernel (csrc/knn_tma.cu) and the tensor-core-filtered kNN (csrc/knn_tc.cu + flagged fall-back)
at the shapes of the bench step; prints times, identity of the graphs, the share of flagged rows and the list statistics.
usage: python tools/exp_knn_tc.py [B] [data]   data: randn | feat (correlated features with a common offset, like a BN + LeakyReLU output)"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "parsenet-codebase_b200"))
import torch
from pnb200 import ops

B = int(sys.argv[1]) if len(sys.argv) > 1 else 16
data = sys.argv[2] if len(sys.argv) > 2 else "feat"
torch.manual_seed(0)


def features(N, C):
    if data == "randn":
        return torch.randn(B, N, C, device="cuda") * 0.3
    # points on a few smooth patches pushed through a random 2-layer map: clustered, correlated channels, non-zero mean
    u = torch.rand(B, N, 3, device="cuda")
    W1 = torch.randn(3, 32, device="cuda"); W2 = torch.randn(32, C, device="cuda") * 0.3
    h = torch.nn.functional.leaky_relu(torch.sin(3 * u @ W1), 0.2)
    return (torch.nn.functional.leaky_relu(h @ W2 + 0.5, 0.2)).contiguous()


def timed(fn, reps=4):
    best = 1e9
    for it in range(reps):
        a = torch.cuda.Event(enable_timing=True); b = torch.cuda.Event(enable_timing=True)
        a.record(); out = fn(); b.record(); torch.cuda.synchronize()
        if it:
            best = min(best, a.elapsed_time(b))
    return out, best


print(f"data = {data}, B = {B}")
print("| N | C | k | exact kernel ms | bracketed ms | speed-up | identical idx | identical dist | flagged rows | plan | mean / max list | ")
print("|---|---|---|---:|---:|---:|---|---|---:|---|---|")
for N, C, k, metric in ((10000, 64, 80, 0), (5000, 64, 10, 0), (5000, 128, 10, 0), (5000, 256, 10, 0), (10000, 6, 80, 1), (5000, 3, 10, 0)):
    if C <= 6:
        x = torch.randn(B, N, C, device="cuda") * 0.3
        if metric == 1:
            x[..., 3:] = torch.nn.functional.normalize(x[..., 3:], dim=-1)
    else:
        x = features(N, C)
    ops.KNN_IMPL = "tma"
    (i0, d0), t0 = timed(lambda: ops.knn_graph(x, k, metric, return_dist=True))
    ops.KNN_IMPL = "tc"
    (i1, d1), t1 = timed(lambda: ops.knn_graph(x, k, metric, return_dist=True))
    plan = ops.knn_tc_plan(N, k, single_list=C <= 6)
    w = [v for kk, v in ops._KNN_WS.items() if kk[2] == plan[2]][0]
    flagged = w["flags"][: B * N].float().mean().item()
    cnt = w["cnt"][: B * N].sum(1).float()
    print(f"| {N} | {C} | {k} | {t0:.2f} | {t1:.2f} | {t0 / t1:.2f}x | {torch.equal(i0, i1)} | {torch.equal(d0, d1)} | {flagged:.5f} | "
          f"{plan} | {cnt.mean().item():.0f} / {cnt.max().item():.0f} |", flush=True)
