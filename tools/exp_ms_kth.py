"""A/B of the K-th distance (mean-shift bandwidth) at the bench size: four-pass radix kernel vs the one-pass bracketed path.
usage: python tools/exp_ms_kth.py [B] [N] [K]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "parsenet-codebase_b200"))
import torch

from pnb200 import meanshift as pms
from pnb200.cabi import call

B = int(sys.argv[1]) if len(sys.argv) > 1 else 16
N = int(sys.argv[2]) if len(sys.argv) > 2 else 10000
K = int(sys.argv[3]) if len(sys.argv) > 3 else 250
g = torch.Generator().manual_seed(0)
cen = torch.nn.functional.normalize(torch.randn(8, 128, generator=g), dim=1)
lab = torch.randint(0, 8, (B, N), generator=g)
X = torch.nn.functional.normalize(cen[lab] + 0.05 * torch.randn(B, N, 128, generator=g), dim=2).cuda().contiguous()


def timed(fn, reps=5):
    for _ in range(2):
        out = fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        out = fn()
    e1.record(); torch.cuda.synchronize()
    return out, e0.elapsed_time(e1) / reps


def radix():
    kth = torch.empty(B, N, device="cuda")
    call("pn_ms_kth_dist_tc", X.data_ptr(), None, B, N, N * 128, 128, K, kth.data_ptr(), torch.cuda.current_stream().cuda_stream)
    return kth


ref, t_radix = timed(radix)
pms.KTH_BRACKET = True
got, t_br = timed(lambda: pms._kth_all_rows(X, K))
print(f"B={B} N={N} K={K} plan (stride, b, cap) = {pms.kth_bracket_plan(N, K)}")
print(f"| radix 4-pass ms | bracketed ms | speed-up | max abs diff | identical rows |")
print(f"| {t_radix:.3f} | {t_br:.3f} | {t_radix / t_br:.2f}x | {(got - ref).abs().max().item():.2e} | {(got == ref).float().mean().item():.5f} |")
