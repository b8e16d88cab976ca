"""profiling driver: a few kNN launches at the BASELINE shape (run under ncu)"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "parsenet-codebase_b200"))
import torch
from pnb200 import ops
B = int(sys.argv[1]) if len(sys.argv) > 1 else 4
torch.manual_seed(0)
x64 = torch.randn(B, 10000, 64, device="cuda") * 0.3
x6 = torch.randn(B, 10000, 6, device="cuda") * 0.3
for _ in range(2):
    ops.knn_graph(x6, 80, 1)
    ops.knn_graph(x64, 80, 0)
torch.cuda.synchronize()
a = torch.cuda.Event(enable_timing=True); b = torch.cuda.Event(enable_timing=True); c = torch.cuda.Event(enable_timing=True)
a.record(); ops.knn_graph(x6, 80, 1); b.record(); ops.knn_graph(x64, 80, 0); c.record(); torch.cuda.synchronize()
print("metric1 ms", a.elapsed_time(b), "c64 ms", b.elapsed_time(c))
