"""bring-up check for the tcgen05 mean-shift backward kernels: compare with the fp32 FMA kernels and time both"""
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
st = torch.cuda.current_stream().cuda_stream
Yn = torch.empty_like(X); den = torch.empty(B, N, device="cuda"); un = torch.empty(B, N, device="cuda")
call("pn_ms_iter_fwd", Y.data_ptr(), X.data_ptr(), B, N, 128, cinv.data_ptr(), Yn.data_ptr(), den.data_ptr(), un.data_ptr(), st)
g = torch.randn_like(X)
res = {}
import threading, time
prog = torch.full((16,), -1, dtype=torch.int32).pin_memory()
call("pn_debug_set_progress", prog.data_ptr())
def watchdog():
    time.sleep(12)
    print("WATCHDOG progress words (slot: epilogue warps 0-7, loaders 8-11, mma-g1 12, mma-g2 13, mma-start 14):", prog.tolist(), flush=True)
threading.Thread(target=watchdog, daemon=True).start()
for name in ("pn_ms_iter_bwd", "pn_ms_iter_bwd_tc"):
    print("running", name, flush=True)
    Gn = torch.empty_like(X); gd = torch.empty(B, N, device="cuda")
    gY = torch.empty_like(X); gX = torch.zeros_like(X)
    args = (g.data_ptr(), Yn.data_ptr(), Y.data_ptr(), X.data_ptr(), den.data_ptr(), un.data_ptr(), B, N, 128,
            cinv.data_ptr(), Gn.data_ptr(), gd.data_ptr(), gY.data_ptr(), gX.data_ptr(), 1, st)
    call(name, *args)
    torch.cuda.synchronize()
    r_gY, r_gX = gY.clone(), gX.clone()
    a = torch.cuda.Event(enable_timing=True); b = torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(2):
        call(name, *args)
    b.record(); torch.cuda.synchronize()
    res[name] = (r_gY, r_gX, a.elapsed_time(b) / 2)
r, t = res["pn_ms_iter_bwd"], res["pn_ms_iter_bwd_tc"]
for i, nm in enumerate(["gY", "gX"]):
    err = (r[i] - t[i]).abs().max().item() / (r[i].abs().max().item() + 1e-30)
    print(f"{nm}: rel err tc vs simt = {err:.3e}")
print(f"B={B} N={N}: simt {r[2]:.3f} ms, tc {t[2]:.3f} ms, speedup {r[2]/t[2]:.2f}x")
