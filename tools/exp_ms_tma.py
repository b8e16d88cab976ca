"""round-2 experiment: TMA-fed operand tiles for the tcgen05 mean-shift kernels (csrc/meanshift_tma.cu, opt-in PN_MS_TMA=1).
A/B inside one process against the default kernels: results should be BIT-IDENTICAL (same products in the same order;
raw fp32 data acts as the truncated "big" part) and the time lower (no loader warps).  The experimental kernels trap
after ~2 s instead of hanging, still: run under `timeout 300`.

    python tools/exp_ms_tma.py            # (2, 1000) then (16, 10000)
    python tools/exp_ms_tma.py 4 4999     # one size
    PN_EXP_CGS=1 python tools/exp_ms_tma.py      # only the 1-CTA kernels (bring those up first; default "1,2":
                                                 # CTA group 2 = CTA pairs, tcgen05 cta_group::2, half the operand traffic)
"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "parsenet-codebase_b200"))
import torch
from pnb200.cabi import call

d = 128
CGS = [int(c) for c in os.environ.get("PN_EXP_CGS", "1,2").split(",")]
st = torch.cuda.current_stream().cuda_stream


def timed(fn, reps=3):
    fn(); torch.cuda.synchronize()
    best = 1e9
    for _ in range(reps):
        a = torch.cuda.Event(enable_timing=True); b = torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize()
        best = min(best, a.elapsed_time(b))
    return best


def report(name, ref, out, t_ref, t_new):
    same = all(torch.equal(a, b) for a, b in zip(ref, out))
    err = max(((a - b).abs().max() / (a.abs().max() + 1e-30)).item() for a, b in zip(ref, out))
    print(f"{name}: default {t_ref:.3f} ms, TMA {t_new:.3f} ms ({t_ref / t_new:.2f}x), bit-identical={same}, "
          f"max rel diff {err:.3e}", flush=True)


def run(B, N):
    torch.manual_seed(0)
    X = torch.nn.functional.normalize(torch.randn(B, N, d, device="cuda"), dim=2)
    Y = torch.nn.functional.normalize(X + 0.05 * torch.randn_like(X), dim=2)
    cinv = torch.tensor([1 / 0.3 ** 2, 1 / 0.8 ** 2] * (B // 2 + 1), device="cuda")[:B].contiguous()
    Np = (N + 31) // 32 * 32
    Xs = torch.empty_like(X); Xt = torch.empty(B, d, Np, device="cuda"); Xst = torch.empty(B, d, Np, device="cuda")
    t_prep = timed(lambda: call("pn_ms_prepare_operands", X.data_ptr(), B, N, d, Np, Xs.data_ptr(), Xt.data_ptr(),
                                Xst.data_ptr(), st))
    hi = (X.view(torch.int32) & -8192).view(torch.float32)
    assert torch.equal(Xs, X - hi) and torch.equal(Xt[:, :, :N], X.transpose(1, 2)) and \
        torch.equal(Xst[:, :, :N], (X - hi).transpose(1, 2)) and (Xt[:, :, N:] == 0).all(), "operand forms wrong"
    print(f"B={B} N={N}: operand forms ok, prep {t_prep:.3f} ms", flush=True)

    def fwd(tma):
        Yn = torch.empty_like(X); den = torch.empty(B, N, device="cuda"); un = torch.empty(B, N, device="cuda")
        if tma:
            f = lambda: call("pn_ms_iter_fwd_tma", Y.data_ptr(), X.data_ptr(), Xs.data_ptr(), Xt.data_ptr(), Xst.data_ptr(),
                             B, N, d, Np, cinv.data_ptr(), Yn.data_ptr(), den.data_ptr(), un.data_ptr(), st)
        else:
            f = lambda: call("pn_ms_iter_fwd_tc", Y.data_ptr(), X.data_ptr(), B, N, d, cinv.data_ptr(), Yn.data_ptr(),
                             den.data_ptr(), un.data_ptr(), st)
        return (Yn, den, un), timed(f)

    ref, t0 = fwd(False)
    for cg in CGS:
        os.environ["PN_MS_TMA_CG"] = str(cg)
        out, t1 = fwd(True)
        report(f"forward, CTA group {cg}", ref, out, t0, t1)
    Yn, den, un = ref
    g = torch.randn_like(X)

    def bwd(tma):
        Gn = torch.empty_like(X); gd = torch.empty(B, N, device="cuda"); gY = torch.empty_like(X); gX = torch.zeros_like(X)
        wsC = torch.empty(4, B, 2 * ((N + 15) // 16 * 16), d, device="cuda")
        if tma:
            f = lambda: call("pn_ms_iter_bwd_tma", g.data_ptr(), Yn.data_ptr(), Y.data_ptr(), X.data_ptr(), Xs.data_ptr(),
                             Xt.data_ptr(), Xst.data_ptr(), den.data_ptr(), un.data_ptr(), B, N, d, Np, cinv.data_ptr(),
                             Gn.data_ptr(), gd.data_ptr(), wsC.data_ptr(), gY.data_ptr(), gX.data_ptr(), 0, st)
        else:
            f = lambda: call("pn_ms_iter_bwd_tc", g.data_ptr(), Yn.data_ptr(), Y.data_ptr(), X.data_ptr(), den.data_ptr(),
                             un.data_ptr(), B, N, d, cinv.data_ptr(), Gn.data_ptr(), gd.data_ptr(), gY.data_ptr(),
                             gX.data_ptr(), 0, st)
        return (gY, gX), timed(f)

    ref, t0 = bwd(False)
    for cg in CGS:
        os.environ["PN_MS_TMA_CG"] = str(cg)
        out, t1 = bwd(True)
        report(f"backward (prep + rows + cols), CTA group {cg}", ref, out, t0, t1)


if len(sys.argv) > 2:
    run(int(sys.argv[1]), int(sys.argv[2]))
else:
    run(2, 1000)
    run(16, 10000)
