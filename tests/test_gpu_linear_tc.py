"""tcgen05 (split-TF32) per-point linear layer vs the fp32 FMA-pipe kernel of the same C-ABI contract and vs a
float64 torch reference of Y = act(norm(A)) W^T + bias + sbias with per-(shape, group) sum / sum-of-squares."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def _run(name, A, W, bias, sbias, sc, sh, act, G):
    from pnb200.cabi import call
    B, Np, K = A.shape
    Nout = W.shape[0]
    Y = torch.full((B, Np, Nout), float("nan"), device="cuda")
    stats = torch.zeros((B, G, 2), dtype=torch.float64, device="cuda") if G else None
    p = lambda t: None if t is None else t.data_ptr()
    call(name, p(A), A.stride(1), p(W), W.stride(0), p(bias), p(sbias), p(sc), p(sh), act, p(Y), Nout, p(stats), B, Np,
         K, Nout, max(G, 1), 1, torch.cuda.current_stream().cuda_stream)
    return Y, stats


@pytest.mark.parametrize("B,Np,K,Nout,G,act,norm", [
    (1, 128, 64, 128, 2, 0, False),        # one full tile
    (2, 300, 256, 1024, 8, 1, True),       # mlp1 shape (ragged rows), ReLU(GroupNorm(.)) on load
    (3, 1000, 512, 256, 4, 1, True),       # head conv2
    (2, 777, 256, 128, 0, 2, True),        # embedding layer, LeakyReLU, no statistics
    (1, 2113, 64, 256, 2, 0, False),       # edge-conv P/Q GEMM (K = 64)
    (2, 129, 20, 32, 1, 0, False),         # K not a multiple of the 16-wide stage, one column chunk
])
def test_linear_fwd_tc_matches_simt_and_fp64(B, Np, K, Nout, G, act, norm):
    g = torch.Generator().manual_seed(B * 1000 + Np)
    A = torch.randn(B, Np, K, generator=g).cuda()
    W = (torch.randn(Nout, K, generator=g) / K ** 0.5).cuda()
    bias = torch.randn(Nout, generator=g).cuda()
    sbias = torch.randn(B, Nout, generator=g).cuda()
    sc = (torch.rand(B, K, generator=g) + 0.5).cuda() if norm else None
    sh = torch.randn(B, K, generator=g).cuda() if norm else None
    Ys, Ss = _run("pn_linear_fwd", A, W, bias, sbias, sc, sh, act, G)
    Yt, St = _run("pn_linear_fwd_tc", A, W, bias, sbias, sc, sh, act, G)
    X = A.double()
    if norm:
        X = X * sc.double().unsqueeze(1) + sh.double().unsqueeze(1)
    X = torch.relu(X) if act == 1 else (torch.where(X > 0, X, 0.2 * X) if act == 2 else X)
    Yr = X @ W.double().t() + bias.double() + sbias.double().unsqueeze(1)
    assert torch.isfinite(Yt).all()
    scale = Yr.abs().max().item()
    err_t = (Yt.double() - Yr).abs().max().item() / scale
    err_s = (Ys.double() - Yr).abs().max().item() / scale
    print(f"rel err vs fp64: tc {err_t:.2e}, fp32 pipe {err_s:.2e}")
    # split-TF32 keeps ~21 mantissa bits per product; the tensor core accumulates with truncation, which adds a bias of
    # ~0.5 ulp per accumulation step (K/8 steps): 3e-6 at K = 512.  Far inside the 1e-4 budget of the north star.
    assert err_t < 1e-5, err_t
    if G:
        cpg = Nout // G
        ref = torch.stack([Yr.view(B, Np, G, cpg).sum((1, 3)), (Yr ** 2).view(B, Np, G, cpg).sum((1, 3))], 2)
        # sums are compared against the magnitude they are made of (sum |y|, sum y^2): a group sum can cancel to ~0
        mag = torch.stack([Yr.abs().view(B, Np, G, cpg).sum((1, 3)), ref[:, :, 1]], 2)
        assert ((St - ref).abs() / mag).max().item() < 1e-5
        assert ((Ss - ref).abs() / mag).max().item() < 1e-5


def test_linear_dispatch_uses_tensor_cores_for_mlp_shapes():
    """ops.linear_fwd routes the dense MLP shapes to pn_linear_fwd_tc and the tiny ones to the FP32-pipe kernel"""
    from pnb200 import cabi, ops
    A = torch.randn(2, 256, 256, device="cuda")
    W = torch.randn(512, 256, device="cuda")
    cabi.TIMED["pn_linear_fwd_tc"] = []
    cabi.TIMED["pn_linear_fwd"] = []
    ops.linear_fwd(A, W, stats_groups=8)
    ops.linear_fwd(torch.randn(2, 256, 6, device="cuda"), torch.randn(128, 6, device="cuda"))
    torch.cuda.synchronize()
    n_tc, n_simt = len(cabi.TIMED.pop("pn_linear_fwd_tc")), len(cabi.TIMED.pop("pn_linear_fwd"))
    assert (n_tc, n_simt) == ((1, 1) if ops.LINEAR_IMPL == "tc" else (0, 2))


@pytest.mark.parametrize("B,Np,K,Nout,G,act,fin,acc", [
    (2, 300, 256, 1024, 8, 1, True, False),      # mlp1: dZ over the 256-wide concat, GroupNorm(8)... of a 256-wide producer
    (3, 1000, 512, 256, 8, 1, True, False),      # head conv2 <- conv1 (GroupNorm(8, 512), ReLU)
    (2, 777, 256, 128, 4, 1, True, False),       # embedding layer
    (1, 2113, 64, 128, 0, 0, False, True),       # edge-conv P/Q GEMM: accumulate into a concat-gradient slice, no finalize
    (2, 500, 128, 256, 0, 2, True, True),        # mask only (no norm sums) + accumulate
])
def test_linear_bwd_data_tc_matches_simt_and_fp64(B, Np, K, Nout, G, act, fin, acc):
    """dZ = dY W with the finalize epilogue (activation mask of the layer input, norm-backward sums) on the tensor cores
    against the FP32-pipe kernel of the same contract and a float64 evaluation"""
    from pnb200.cabi import call
    g = torch.Generator().manual_seed(B * 100 + Np)
    dY = torch.randn(B, Np, Nout, generator=g).cuda()
    W = (torch.randn(Nout, K, generator=g) / Nout ** 0.5).cuda()
    A = torch.randn(B, Np, K, generator=g).cuda()
    sc = (torch.rand(B, K, generator=g) + 0.5).cuda() if fin else None
    sh = torch.randn(B, K, generator=g).cuda() if fin else None
    gamma = (torch.randn(K, generator=g) * 0.5 + 1).cuda() if (fin and G) else None
    mr = torch.stack([torch.randn(B, max(G, 1), generator=g) * 0.1, torch.rand(B, max(G, 1), generator=g) + 0.5], 2).cuda()
    base = torch.randn(B, Np, K, generator=g).cuda()
    st = torch.cuda.current_stream().cuda_stream
    p = lambda t: None if t is None else t.data_ptr()
    outs = []
    for name, Wop in (("pn_linear_bwd_data", W), ("pn_linear_bwd_data_tc", W.t().contiguous())):
        dZ = base.clone() if acc else torch.full((B, Np, K), float("nan"), device="cuda")
        gsum = torch.zeros(B, max(G, 1), 2, dtype=torch.float64, device="cuda") if gamma is not None else None
        call(name, p(dY), Nout, p(Wop), Wop.stride(0), p(dZ), K, 1 if acc else 0, 1 if fin else 0, p(A) if fin else None,
             K if fin else 0, p(sc), p(sh), act, p(gamma), p(mr) if gamma is not None else None, p(gsum), B, Np, K, Nout,
             max(G, 1), 1, st)
        outs.append((dZ, gsum))
    ref = dY.double() @ W.double()
    if acc:
        ref = ref + base.double()
    if fin:
        pre = A.double() * sc.double().unsqueeze(1) + sh.double().unsqueeze(1)
        mask = (pre > 0).double() if act == 1 else (torch.where(pre > 0, 1.0, 0.2) if act == 2 else torch.ones_like(pre))
        ref = ref * mask
    scale = ref.abs().max().item()
    err_s = (outs[0][0].double() - ref).abs().max().item() / scale
    err_t = (outs[1][0].double() - ref).abs().max().item() / scale
    print(f"rel err vs fp64: tc {err_t:.2e}, fp32 pipe {err_s:.2e}")
    assert torch.isfinite(outs[1][0]).all() and err_t < 1e-5, err_t
    if gamma is not None:
        cpg = K // G
        xh = (A.double() - mr[:, :, 0].double().repeat_interleave(cpg, 1).unsqueeze(1)) * \
            mr[:, :, 1].double().repeat_interleave(cpg, 1).unsqueeze(1)
        gt = ref * gamma.double()
        want = torch.stack([gt.view(B, Np, G, cpg).sum((1, 3)), (gt * xh).view(B, Np, G, cpg).sum((1, 3))], 2)
        mag = torch.stack([gt.abs().view(B, Np, G, cpg).sum((1, 3)), (gt * xh).abs().view(B, Np, G, cpg).sum((1, 3))], 2)
        for o in outs:
            assert ((o[1] - want).abs() / mag).max().item() < 1e-5


@pytest.mark.parametrize("B,Np,K,Nout,act,norm,bias", [
    (2, 300, 256, 1024, 1, True, "b"),        # mlp1 (ragged row count, several row splits)
    (3, 1000, 512, 256, 1, True, "b"),        # head conv2
    (2, 777, 1280 - 1024, 512, 0, False, "sb"),   # head conv1 on the 256 local channels, per-shape bias gradient
    (1, 2113, 64, 128, 0, False, ""),         # edge-conv P/Q GEMM (K = 64: half a column tile), no bias
    (2, 4999, 128, 132, 2, True, "b"),        # Nout not a multiple of 32 (ragged last row tile)
])
def test_linear_bwd_weight_tc_matches_simt_and_fp64(B, Np, K, Nout, act, norm, bias):
    """dW += dY^T act(norm(A)) (+ bias / per-shape bias gradients) on the tensor cores (transposing loaders, row splits, fp32
    atomics) against the FP32-pipe kernel of the same contract and a float64 evaluation"""
    from pnb200.cabi import call
    g = torch.Generator().manual_seed(B * 100 + Np)
    dY = torch.randn(B, Np, Nout, generator=g).cuda()
    A = torch.randn(B, Np, K, generator=g).cuda()
    sc = (torch.rand(B, K, generator=g) + 0.5).cuda() if norm else None
    sh = torch.randn(B, K, generator=g).cuda() if norm else None
    st = torch.cuda.current_stream().cuda_stream
    p = lambda t: None if t is None else t.data_ptr()
    outs = []
    for name in ("pn_linear_bwd_weight", "pn_linear_bwd_weight_tc"):
        dW = torch.zeros(Nout, K, device="cuda")
        db = torch.zeros(Nout, device="cuda") if "b" in bias and bias != "sb" else None
        dsb = torch.zeros(B, Nout, device="cuda") if bias == "sb" else None
        call(name, p(dY), Nout, p(A), K, p(sc), p(sh), act, p(dW), K, p(db), p(dsb), B, Np, K, Nout, st)
        outs.append((dW, db, dsb))
    X = A.double()
    if norm:
        X = X * sc.double().unsqueeze(1) + sh.double().unsqueeze(1)
    X = torch.relu(X) if act == 1 else (torch.where(X > 0, X, 0.2 * X) if act == 2 else X)
    ref = torch.einsum("bmn,bmk->nk", dY.double(), X)
    # magnitude the sums are made of (a weight-gradient entry can cancel to ~0)
    mag = torch.einsum("bmn,bmk->nk", dY.double().abs(), X.abs()).max().item()
    err_s = (outs[0][0].double() - ref).abs().max().item() / mag
    err_t = (outs[1][0].double() - ref).abs().max().item() / mag
    print(f"err / magnitude vs fp64: tc {err_t:.2e}, fp32 pipe {err_s:.2e}")
    assert torch.isfinite(outs[1][0]).all() and err_t < 2e-6, err_t
    if outs[1][1] is not None:
        want = dY.double().sum((0, 1))
        assert ((outs[1][1].double() - want).abs().max() / dY.double().abs().sum((0, 1)).max()).item() < 1e-6
    if outs[1][2] is not None:
        want = dY.double().sum(1)
        assert ((outs[1][2].double() - want).abs().max() / dY.double().abs().sum(1).max()).item() < 1e-6
