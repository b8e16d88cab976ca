"""stress / device-index check: every tcgen05 kernel against its FP32-pipe twin on the selected device, repeated"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "parsenet-codebase_b200"))
import torch
dev_index = int(sys.argv[1]) if len(sys.argv) > 1 else 0
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 20
torch.cuda.set_device(dev_index)
from pnb200.cabi import call
st = lambda: torch.cuda.current_stream().cuda_stream
print("device", torch.cuda.current_device(), torch.cuda.get_device_name())
torch.manual_seed(0)
B, Np = 16, 10000
bad = {"linear": 0, "argsel0": 0, "argsel1": 0, "msfwd": 0, "kth": 0}
for rep in range(reps):
    # linear (embedding layer shape 256 -> 128 and mlp1 256 -> 1024)
    for K, Nout, G in ((256, 128, 0), (256, 1024, 8)):
        A = torch.randn(B, Np, K, device="cuda"); W = torch.randn(Nout, K, device="cuda") / 16
        bias = torch.randn(Nout, device="cuda"); sc = torch.rand(B, K, device="cuda") + 0.5; sh = torch.randn(B, K, device="cuda")
        outs = []
        for name in ("pn_linear_fwd", "pn_linear_fwd_tc"):
            Y = torch.full((B, Np, Nout), float("nan"), device="cuda")
            stt = torch.zeros(B, max(G, 1), 2, dtype=torch.float64, device="cuda")
            call(name, A.data_ptr(), K, W.data_ptr(), K, bias.data_ptr(), None, sc.data_ptr(), sh.data_ptr(), 1, Y.data_ptr(),
                 Nout, stt.data_ptr() if G else None, B, Np, K, Nout, max(G, 1), 1, st())
            outs.append(Y)
        err = (outs[0] - outs[1]).abs().max().item() / outs[0].abs().max().item()
        if not (err < 1e-4):
            bad["linear"] += 1
            d = ~torch.isfinite(outs[1]) | ((outs[0] - outs[1]).abs() > 1e-3)
            rows = d.any(2).nonzero()
            print(f"rep {rep} linear {K}->{Nout}: err {err}, bad rows {rows.shape[0]}, first {rows[:4].tolist()}", flush=True)
    # mean-shift forward + arg-selects + kth
    X = torch.nn.functional.normalize(torch.randn(B, Np, 128, device="cuda"), dim=2)
    cinv = torch.full((B,), 1.0 / 0.64, device="cuda")
    res = []
    for name in ("pn_ms_iter_fwd", "pn_ms_iter_fwd_tc"):
        Yn = torch.full_like(X, float("nan")); den = torch.empty(B, Np, device="cuda"); un = torch.empty(B, Np, device="cuda")
        call(name, X.data_ptr(), X.data_ptr(), B, Np, 128, cinv.data_ptr(), Yn.data_ptr(), den.data_ptr(), un.data_ptr(), st())
        res.append(Yn)
    e = (res[0] - res[1]).abs().max().item()
    if not (e < 1e-4):
        bad["msfwd"] += 1; print(f"rep {rep} msfwd err {e}", flush=True)
    Y = res[0]
    cnt = torch.randint(0, 5, (B, Np), device="cuda").float(); thr = torch.full((B,), 0.8, device="cuda")
    for mode in (0, 1):
        o = []
        for name in ("pn_ms_argsel", "pn_ms_argsel_tc"):
            out = torch.full((B, Np), -1, dtype=torch.int32, device="cuda")
            call(name, mode, X.data_ptr(), Np * 128, Np, Y.data_ptr(), Np * 128, Np, B, 128, cnt.data_ptr(), thr.data_ptr(),
                 out.data_ptr(), st())
            o.append(out)
        frac = (o[0] != o[1]).float().mean().item()
        oob = ((o[1] < 0) | (o[1] >= Np)).sum().item()
        if frac > 1e-2 or oob:
            bad[f"argsel{mode}"] += 1; print(f"rep {rep} argsel{mode}: mismatch {frac}, out-of-range {oob}", flush=True)
    k = []
    for name in ("pn_ms_kth_dist", "pn_ms_kth_dist_tc"):
        kth = torch.full((B, Np), float("nan"), device="cuda")
        call(name, X.data_ptr(), None, B, Np, Np * 128, 128, 250, kth.data_ptr(), st())
        k.append(kth)
    e = (k[0] - k[1]).abs().max().item()
    if not (e < 1e-4):
        bad["kth"] += 1; print(f"rep {rep} kth err {e}", flush=True)
torch.cuda.synchronize()
print("failures:", bad)
