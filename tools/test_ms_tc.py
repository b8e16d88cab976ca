"""bring-up check for the tcgen05 mean-shift forward kernel: compare with the fp32 FMA kernel and time both"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "parsenet-codebase_b200"))
import torch
from pnb200.cabi import call
B, N = (int(sys.argv[1]), int(sys.argv[2])) if len(sys.argv) > 2 else (2, 1000)
torch.manual_seed(0)
X = torch.nn.functional.normalize(torch.randn(B, N, 128, device="cuda"), dim=2)
Y = torch.nn.functional.normalize(X + 0.05 * torch.randn_like(X), dim=2)
cinv = torch.tensor([1 / 0.3 ** 2, 1 / 0.8 ** 2] * (B // 2 + 1), device="cuda")[:B].contiguous()
outs = {}
for name in ("pn_ms_iter_fwd", "pn_ms_iter_fwd_tc"):
    Yn = torch.empty_like(X); den = torch.empty(B, N, device="cuda"); un = torch.empty(B, N, device="cuda")
    st = torch.cuda.current_stream().cuda_stream
    call(name, Y.data_ptr(), X.data_ptr(), B, N, 128, cinv.data_ptr(), Yn.data_ptr(), den.data_ptr(), un.data_ptr(), st)
    torch.cuda.synchronize()
    a = torch.cuda.Event(enable_timing=True); b = torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(3):
        call(name, Y.data_ptr(), X.data_ptr(), B, N, 128, cinv.data_ptr(), Yn.data_ptr(), den.data_ptr(), un.data_ptr(), st)
    b.record(); torch.cuda.synchronize()
    outs[name] = (Yn.clone(), den.clone(), un.clone(), a.elapsed_time(b) / 3)
r, t = outs["pn_ms_iter_fwd"], outs["pn_ms_iter_fwd_tc"]
for i, nm in enumerate(["Ynew", "den", "unorm"]):
    err = (r[i] - t[i]).abs().max().item() / (r[i].abs().max().item() + 1e-30)
    print(f"{nm}: rel err tc vs simt = {err:.3e}")
print(f"B={B} N={N}: simt {r[3]:.3f} ms, tc {t[3]:.3f} ms, speedup {r[3]/t[3]:.2f}x")
