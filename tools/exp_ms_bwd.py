"""timing-only ablation sweep of the tcgen05 mean-shift backward kernels (PN_MS_BWD_ABLATE, see csrc/meanshift_tc_bwd.cu):
which part of a tile step (exp epilogue, split stores, transposed copy, P stores, first / second product MMAs) the
kernel time is sensitive to.  Results of the ablated variants are wrong by construction; only variant 0 / 1 are real."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "parsenet-codebase_b200"))
import torch
from pnb200.cabi import call

B, N, d = 16, 10000, 128
torch.manual_seed(0)
X = torch.nn.functional.normalize(torch.randn(B, N, d, device="cuda"), dim=2)
Y = torch.nn.functional.normalize(X + 0.05 * torch.randn_like(X), dim=2)
cinv = torch.full((B,), 1.0 / 0.8 ** 2, device="cuda")
st = torch.cuda.current_stream().cuda_stream
Yn = torch.empty_like(X); den = torch.empty(B, N, device="cuda"); un = torch.empty(B, N, device="cuda")
call("pn_ms_iter_fwd_tc", Y.data_ptr(), X.data_ptr(), B, N, d, cinv.data_ptr(), Yn.data_ptr(), den.data_ptr(), un.data_ptr(), st)
g = torch.randn_like(X)
Gn = torch.empty_like(X); gd = torch.empty(B, N, device="cuda"); gY = torch.empty_like(X); gX = torch.zeros_like(X)
NAMES = {0: "exact (3+3 MMAs)", 1: "LITE (P rounded, 3+2)", 2: "no exp", 4: "no small-part stores", 8: "no 2nd-product MMAs",
         16: "1 of 3 MMAs in 1st product", 32: "no transposed copy", 64: "no P stores", 36: "no small + no transposed",
         24: "1of3 + no 2nd product", 126: "all ablations", 128: "double-buffered P (default)", 129: "double-buffered P + LITE",
         254: "double-buffered P + all ablations"}
VARS = [int(v) for v in sys.argv[1].split(",")] if len(sys.argv) > 1 else (0, 1, 2, 4, 32, 36, 64, 8, 16, 24, 126, 128, 129, 254)


def run(var, only):
    os.environ["PN_MS_BWD_ABLATE"] = str(var)
    os.environ["PN_MS_BWD_TC_ONLY"] = only
    best = 1e9
    for rep in range(3):
        a = torch.cuda.Event(enable_timing=True); b = torch.cuda.Event(enable_timing=True)
        a.record()
        call("pn_ms_iter_bwd_tc", g.data_ptr(), Yn.data_ptr(), Y.data_ptr(), X.data_ptr(), den.data_ptr(), un.data_ptr(), B, N, d,
             cinv.data_ptr(), Gn.data_ptr(), gd.data_ptr(), gY.data_ptr(), gX.data_ptr(), 0, st)
        b.record(); torch.cuda.synchronize()
        best = min(best, a.elapsed_time(b))
    return best


print("| variant | rows ms | cols ms |\n|---|---:|---:|")
for var in VARS:
    print(f"| {var}: {NAMES[var]} | {run(var, 'r'):.2f} | {run(var, 'c'):.2f} |", flush=True)
