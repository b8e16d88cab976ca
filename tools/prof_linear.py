"""profiling driver: the dense per-point MLP GEMMs at the BASELINE shape, FP32-pipe kernel vs tcgen05 kernel"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "parsenet-codebase_b200"))
import torch
from pnb200.cabi import call
B, Np = int(sys.argv[1]) if len(sys.argv) > 1 else 16, 10000
torch.manual_seed(0)
for K, Nout, G in ((256, 1024, 8), (256, 512, 8), (512, 256, 4), (256, 256, 4), (256, 128, 0), (64, 256, 0)):
    A = torch.randn(B, Np, K, device="cuda"); W = torch.randn(Nout, K, device="cuda") / K ** 0.5
    bias = torch.randn(Nout, device="cuda"); sc = torch.rand(B, K, device="cuda") + 0.5; sh = torch.randn(B, K, device="cuda")
    Y = torch.empty(B, Np, Nout, device="cuda")
    st = torch.zeros(B, max(G, 1), 2, dtype=torch.float64, device="cuda")
    res = {}
    for name in ("pn_linear_fwd", "pn_linear_fwd_tc"):
        args = (A.data_ptr(), K, W.data_ptr(), K, bias.data_ptr(), None, sc.data_ptr(), sh.data_ptr(), 1, Y.data_ptr(),
                Nout, st.data_ptr() if G else None, B, Np, K, Nout, max(G, 1), 1, torch.cuda.current_stream().cuda_stream)
        for _ in range(3):
            call(name, *args)
        a = torch.cuda.Event(enable_timing=True); b = torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(10):
            call(name, *args)
        b.record(); torch.cuda.synchronize()
        res[name] = a.elapsed_time(b) / 10
    fl = 2.0 * B * Np * K * Nout
    print(f"K={K:4d} Nout={Nout:4d}: fp32 pipe {res['pn_linear_fwd']:.3f} ms ({fl / res['pn_linear_fwd'] / 1e9:.1f} TFLOP/s)  "
          f"tcgen05 {res['pn_linear_fwd_tc']:.3f} ms ({fl / res['pn_linear_fwd_tc'] / 1e9:.1f} TFLOP/s fp32-accurate)")
