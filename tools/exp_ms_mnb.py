"""round-2 experiment: MN-major B operand for the second product of the tcgen05 mean-shift kernels (no transposed XB copy).
For the forward kernel (PN_MS_FWD_MNB=1|2) and the backward kernels (PN_MS_BWD_ABLATE=384|896) prints whether the result
is BIT-IDENTICAL to the default kernels (same products, same order) and the time of each.  Adopt the variant that is
identical; if neither is, the descriptor convention is something else (see tools/tc_probe/probe.cu modes 2, 3).
Run under `timeout 120` -- a wrong descriptor cannot hang (it only reads other shared-memory bytes), but be safe."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "parsenet-codebase_b200"))
import torch
from pnb200.cabi import call

B, N, d = (int(sys.argv[1]), int(sys.argv[2]), 128) if len(sys.argv) > 2 else (16, 10000, 128)
torch.manual_seed(0)
X = torch.nn.functional.normalize(torch.randn(B, N, d, device="cuda"), dim=2)
Y = torch.nn.functional.normalize(X + 0.05 * torch.randn_like(X), dim=2)
cinv = torch.full((B,), 1.0 / 0.8 ** 2, device="cuda")
st = torch.cuda.current_stream().cuda_stream


def timed(fn, reps=3):
    fn(); torch.cuda.synchronize()
    best = 1e9
    for _ in range(reps):
        a = torch.cuda.Event(enable_timing=True); b = torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize()
        best = min(best, a.elapsed_time(b))
    return best


def fwd(mnb):
    os.environ["PN_MS_FWD_MNB"] = str(mnb)
    Yn = torch.empty_like(X); den = torch.empty(B, N, device="cuda"); un = torch.empty(B, N, device="cuda")
    ms = timed(lambda: call("pn_ms_iter_fwd_tc", Y.data_ptr(), X.data_ptr(), B, N, d, cinv.data_ptr(), Yn.data_ptr(),
                            den.data_ptr(), un.data_ptr(), st))
    return (Yn, den, un), ms


ref, t0 = fwd(0)
print(f"fwd default: {t0:.3f} ms")
for m in (1, 2):
    out, t = fwd(m)
    same = all(torch.equal(a, b) for a, b in zip(ref, out))
    err = max(((a - b).abs().max() / (a.abs().max() + 1e-30)).item() for a, b in zip(ref, out))
    print(f"fwd PN_MS_FWD_MNB={m}: {t:.3f} ms, bit-identical={same}, max rel diff {err:.3e}")
os.environ["PN_MS_FWD_MNB"] = "0"
Yn, den, un = ref
g = torch.randn_like(X)


def bwd(var):
    os.environ["PN_MS_BWD_ABLATE"] = str(var)
    Gn = torch.empty_like(X); gd = torch.empty(B, N, device="cuda"); gY = torch.empty_like(X); gX = torch.zeros_like(X)
    ms = timed(lambda: call("pn_ms_iter_bwd_tc", g.data_ptr(), Yn.data_ptr(), Y.data_ptr(), X.data_ptr(), den.data_ptr(),
                            un.data_ptr(), B, N, d, cinv.data_ptr(), Gn.data_ptr(), gd.data_ptr(), gY.data_ptr(),
                            gX.data_ptr(), 0, st))
    return (gY, gX), ms


ref, t0 = bwd(128)
print(f"bwd default (128): {t0:.3f} ms (prep + rows + cols)")
for v in (384, 896, 160):
    out, t = bwd(v)
    same = all(torch.equal(a, b) for a, b in zip(ref, out))
    err = max(((a - b).abs().max() / (a.abs().max() + 1e-30)).item() for a, b in zip(ref, out))
    note = " (timing-only ablation: wrong by construction)" if v == 160 else ""
    print(f"bwd PN_MS_BWD_ABLATE={v}: {t:.3f} ms, bit-identical={same}, max rel diff {err:.3e}{note}")
