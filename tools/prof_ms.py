"""profiling driver: mean-shift iteration forward + backward at the BASELINE shape (run under ncu)"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "parsenet-codebase_b200"))
import torch
from pnb200 import meanshift as pms
B = int(sys.argv[1]) if len(sys.argv) > 1 else 16
its = int(sys.argv[2]) if len(sys.argv) > 2 else 2
torch.manual_seed(0)
X = torch.nn.functional.normalize(torch.randn(B, 10000, 128, device="cuda"), dim=2).requires_grad_()
bw = torch.full((B,), 0.8, device="cuda")
for rep in range(2):
    a = torch.cuda.Event(enable_timing=True); b = torch.cuda.Event(enable_timing=True); c = torch.cuda.Event(enable_timing=True)
    a.record()
    Y = pms.mean_shift_iters(X, bw, its)
    b.record()
    Y.sum().backward()
    c.record(); torch.cuda.synchronize()
    print(f"B={B} its={its}: fwd {a.elapsed_time(b)/its:.2f} ms/it, bwd {b.elapsed_time(c)/its:.2f} ms/it")
